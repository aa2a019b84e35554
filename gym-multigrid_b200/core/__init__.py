"""Constant tables of `gym_multigrid.core` that user code reads (action enums, world index tables); the objects, grids and agents
themselves live on the device as state planes."""
