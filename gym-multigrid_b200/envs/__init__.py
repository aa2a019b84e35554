"""The reference's env modules under this package's name (`gym_multigrid.envs.{collect_game,ctf,maze}`): each name resolves to
the single-env adaptor over the CUDA classes.  Imported lazily - the classes need torch and a B200."""
