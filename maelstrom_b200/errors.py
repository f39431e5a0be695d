"""Maelstrom's error-code registry (restated from resources/errors.edn:2-44); used by the
host-side client mirror exactly like maelstrom.client/error-registry (client.clj:22-39)."""

ERRORS = {
    0: ("timeout", False),
    1: ("node-not-found", True),
    10: ("not-supported", True),
    11: ("temporarily-unavailable", True),
    12: ("malformed-request", True),
    13: ("crash", False),
    14: ("abort", True),
    20: ("key-does-not-exist", True),
    21: ("key-already-exists", True),
    22: ("precondition-failed", True),
    30: ("txn-conflict", True),
}


def name(code):
    return ERRORS.get(code, ("unknown", False))[0]


def definite(code):
    return ERRORS.get(code, ("unknown", False))[1]
