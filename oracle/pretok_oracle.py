"""Opt-in regex pre-tokenisation -- TEST INFRASTRUCTURE, restated from reference shredword/base.py:38-58.

apply_regex: the reference compiles ONE pattern (its "regex_pattern1", the GPT-4 style split) with the `regex` module and
returns regex.findall(pattern, text). Restated here with the same pattern string; pinned by tests/golden/pretok_cases.json,
which holds the outputs of the reference's OWN apply_regex (imported from /root/reference in the build container by
tests/golden/make_pretok_golden.py).

pretokenize_bytes: how the product feeds the pieces to the whitespace-splitting trainer / encoder -- every piece followed
by one ' ', the trainer's four delimiter bytes INSIDE a piece remapped to the ASCII separators 0x1C-0x1F (which the
pattern's \\s does not match and text does not contain). The pieces cover the text, so undo_pretokenize restores it."""
import regex

# reference base.py:55
PATTERN = r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+"""
_compiled = regex.compile(PATTERN)
_REMAP = bytes.maketrans(b" \t\n\r", b"\x1c\x1d\x1e\x1f")
_UNMAP = bytes.maketrans(b"\x1c\x1d\x1e\x1f", b" \t\n\r")


def apply_regex(text: str) -> list:
  return regex.findall(_compiled, text)  # reference base.py:56-58


def pretokenize_bytes(data: bytes) -> bytes:
  pieces = apply_regex(bytes(data).decode("utf-8"))
  return b"".join(p.encode("utf-8").translate(_REMAP) + b" " for p in pieces)


def undo_pretokenize(data: bytes) -> bytes:
  return bytes(data).replace(b" ", b"").translate(_UNMAP)
