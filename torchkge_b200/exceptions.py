"""Exception classes with the reference's names (torchkge/exceptions.py:8-45), so that
``except torchkge.exceptions.NotYetEvaluatedError`` style handlers keep working after the
import is switched to this package."""


class NotYetEvaluatedError(Exception):
    pass


class SizeMismatchError(Exception):
    pass


class WrongDimensionError(Exception):
    pass


class NotYetImplementedError(Exception):
    pass


class WrongArgumentsError(Exception):
    pass


class SanityError(Exception):
    pass


class SplitabilityError(Exception):
    pass


class NoPreTrainedVersionError(Exception):
    pass
