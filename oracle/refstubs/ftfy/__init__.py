"""Stand-in for `ftfy` (reference utils.py:12): identity is exact for the ASCII class names used here."""


def fix_text(t):
    return t
