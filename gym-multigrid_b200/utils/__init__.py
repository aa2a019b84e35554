"""Host-side helpers user code imports from `gym_multigrid.utils` (map loader, distances); numpy only."""
