"""Header sanitizer used when tabular headers become `extras` keys
(/root/reference/src/itaxotools/taxi2/encoding.py:89-93): NFKC-normalise, drop leading
punctuation, then collapse every run of characters that is neither a word character nor a space
into one underscore.

Out of scope ("Id sanitizer"): the reference's extended-ASCII transliteration
table (e.g. u-umlaut -> "ue").  Headers made of ASCII word characters, the only kind in the
shipped samples, sanitize identically.
"""
from __future__ import annotations

import re
import unicodedata

_LEADING = re.compile(r"^[^\w ]+")
_RUN = re.compile(r"[^\w ]+")


def sanitize(text: str) -> str:
    text = unicodedata.normalize("NFKC", text)
    return _RUN.sub("_", _LEADING.sub("", text))
