"""Exception and warning names of hop/exceptions.py, kept for API compatibility."""


class MissingDataError(Exception):
    """Signifies a critical error because required data is missing."""


class SelectionError(Exception):
    """Signifies an error with a atom selection."""


class MissingDataWarning(Warning):
    """Warns when prerequisite data/files are missing."""


class InconsistentDataWarning(Warning):
    """Warns when the user provides input that is inconsistent."""


class MissingTopologyWarning(Warning):
    """Warns when a topology (psf) is not available."""
