"""Error and warning types, with the semantics of the reference's ``pycc/exceptions.py:14-50``:
``InvalidKeywordError`` is both a ``PyCCError`` and a ``ValueError`` and carries
``keyword`` / ``value`` / ``allowed``; ``PyCCWarning`` is a ``UserWarning``."""
from __future__ import annotations


class PyCCError(Exception):
    pass


class PyCCWarning(UserWarning):
    pass


class InvalidKeywordError(PyCCError, ValueError):
    def __init__(self, keyword, value, allowed):
        self.keyword, self.value, self.allowed = keyword, value, list(allowed)
        choices = ", ".join(repr(x) for x in self.allowed)
        super().__init__("%r is not an allowed value for '%s'. Allowed values: %s."
                         % (value, keyword, choices))
