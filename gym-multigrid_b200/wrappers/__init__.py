from .toroid import ToroidObservation  # noqa: F401
