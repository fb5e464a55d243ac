"""Exception types carrying the reference's names (torchkge/exceptions.py:8-45), so that handlers
written as ``except torchkge.exceptions.NotYetEvaluatedError`` keep working once the import is
switched to this package.  They share one base (``KgeError``) here and say when they are raised."""


class KgeError(Exception):
    """Base of every exception this package raises on behalf of the torchkge API."""


def _named(name, when):
    return type(name, (KgeError,), {"__doc__": when, "__module__": __name__})


NotYetEvaluatedError = _named(
    "NotYetEvaluatedError", "A metric getter of an evaluator was called before evaluate().")
SizeMismatchError = _named(
    "SizeMismatchError", "Tensors that must have equal lengths do not (e.g. heads / tails / relations).")
WrongDimensionError = _named(
    "WrongDimensionError", "An embedding dimension does not fit the model.")
NotYetImplementedError = _named(
    "NotYetImplementedError", "The requested variant exists in torchkge's API but has no implementation.")
WrongArgumentsError = _named(
    "WrongArgumentsError", "An argument is outside its documented set (e.g. missing='heads'|'tails').")
SanityError = _named(
    "SanityError", "A knowledge graph failed a consistency check.")
SplitabilityError = _named(
    "SplitabilityError", "A graph cannot be split as requested without losing entities or relations.")
NoPreTrainedVersionError = _named(
    "NoPreTrainedVersionError", "No pretrained weights exist for the requested model / dataset / dimension.")
