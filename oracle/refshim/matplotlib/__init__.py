"""Empty stand-in so gym_multigrid/utils/window.py:7-12 does not sys.exit (test infra only)."""
