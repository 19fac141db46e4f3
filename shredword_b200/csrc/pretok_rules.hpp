// pretok_rules.hpp -- opt-in regex pre-tokenisation: where the pieces of the reference's split pattern end.
//
// The reference (shredword/base.py:38-58, apply_regex) compiles ONE pattern with the `regex` module and returns
// regex.findall(pattern, text):
//
//   '(?i:[sdmt]|ll|ve|re) | [^\r\n\p{L}\p{N}]?+\p{L}+ | \p{N}{1,3} | ?[^\s\p{L}\p{N}]++[\r\n]* | \s*[\r\n] | \s+(?!\S) | \s+
//        (1)                      (2)                      (3)             (4)                     (5)         (6)        (7)
//
// The pieces cover the text, so the result is fully described by "does a piece end after character k". A regex engine finds
// that left to right; here it is a LOCAL predicate of k, which is what lets every character be decided by its own thread.
// With the classes L = \p{L}, N = \p{N}, S = \s, O = everything else, nl = CR or LF, c = character k, p / f = the characters
// before / after it:
//
//   c in L : f not in L, or c is the last letter of a contraction (1). A match starts at an apostrophe iff p is neither O
//            (then (4) took the whole run of O characters, possessively) nor U+0020 (then (4) started at that space); the
//            letters are tried in the pattern's order: one of [sdmt] first, then ll / ve / re, case-insensitively as the
//            `regex` module folds them (U+017F counts as s).
//   c in N : f not in N, or c is the 3rd, 6th, ... digit of its run (only (3) consumes digits, and it starts with the run).
//   c in O : f in O -> no. f in N, f = end of text -> yes. f in S -> yes unless f is nl ((4) goes on over CR / LF).
//            f in L -> yes iff p in O (the run was longer than c: (4) took it) or p = U+0020 ((4) started there); otherwise
//            c is the prefix character of (2) or the apostrophe of (1).
//   c in S : let R be the maximal run of S characters around c.
//            * nl characters at the start of R directly after an O character belong to that character's piece (4):
//              the last of them ends it.
//            * otherwise no piece ends before the LAST nl of R ((5) is greedy), one ends at it,
//            * and behind it: if R reaches the end of the text it is one piece (6); else a piece ends before the last
//              character of R ((6) gives one back), and after it unless f is a letter (then it is the prefix of (2)) or
//              it is U+0020 and f is in O (then (4) starts with it).
//
// The same functions compile for the host: tests/pretok_host_check.cpp runs them on the CPU against the reference's own
// outputs (tests/golden/pretok_cases.json) -- test infrastructure only; the library exports the device path alone.
//
// Bytes that are not well-formed UTF-8 (the reference only ever sees str) are characters of class O: a character is a
// non-continuation byte with all continuation bytes that follow it.
#pragma once

#include <cstdint>
#include <cstring>

#ifndef SWB_HD
#ifdef __CUDACC__
#define SWB_HD __host__ __device__ __forceinline__
#else
#define SWB_HD inline
#endif
#endif

namespace swb {

#include "unicode_ranges.inc"

constexpr uint32_t PT_TABLE_BYTES = 0x110000 / 4;  // 2 bits per code point
constexpr uint32_t PT_O = 0, PT_L = 1, PT_N = 2, PT_S = 3, PT_NONE = 4;
constexpr uint32_t PT_INVALID = 0xFFFFFFFFu;

// 2-bit class per code point, expanded from the generated ranges (host, once per process)
inline void pretok_build_table(uint8_t *tab) {
  memset(tab, 0, PT_TABLE_BYTES);
  for (const auto &r : SWB_UNICODE_RANGES)
    for (uint32_t cp = r[0]; cp <= r[1]; cp++) tab[cp >> 2] |= (uint8_t)(r[2] << ((cp & 3) * 2));
}

struct PtChar {
  uint32_t cp;   // code point, PT_INVALID for malformed bytes
  uint32_t cls;  // PT_O / PT_L / PT_N / PT_S, PT_NONE outside the text
  uint64_t at;   // index of its first byte
  uint64_t nx;   // index of the next character's first byte
};

#ifndef SWB_TX_NEAR
#define SWB_TX_NEAR
// text[i] for an i within 128 bytes of the byte being decided (a windowed text answers it from its staged copy, unchecked)
template <class T>
SWB_HD uint8_t tx_near(const T &text, uint64_t i) { return text[i]; }
#endif

SWB_HD bool pt_is_cont(uint8_t b) { return (b & 0xC0) == 0x80; }
SWB_HD bool pt_is_nl(const PtChar &c) { return c.cp == '\n' || c.cp == '\r'; }

SWB_HD uint32_t pt_class(const uint8_t *__restrict__ tab, uint32_t cp) {
  if (cp < 128) {
    if ((cp | 32) - 'a' < 26u) return PT_L;
    if (cp - '0' < 10u) return PT_N;
    return (cp - 9 < 5u || cp == 32) ? PT_S : PT_O;
  }
  if (cp >= 0x110000u) return PT_O;
  return (tab[cp >> 2] >> ((cp & 3) * 2)) & 3u;
}

// the character whose first byte is text[i] (i < n; text[i] is not a continuation byte unless i == 0 or the text is malformed)
template <class T>
SWB_HD PtChar pt_at(const T &text, uint64_t n, const uint8_t *__restrict__ tab, uint64_t i) {
  PtChar c;
  c.at = i;
  const uint32_t b0 = text[i];
  uint64_t e = i + 1;
  if (b0 < 0x80) {
    if (e < n && pt_is_cont(text[e])) {  // stray continuation bytes glued to an ASCII byte: malformed
      while (e < n && pt_is_cont(text[e])) ++e;
      c.cp = PT_INVALID; c.cls = PT_O; c.nx = e;
      return c;
    }
    c.cp = b0; c.cls = pt_class(tab, b0); c.nx = e;
    return c;
  }
  while (e < n && pt_is_cont(text[e])) ++e;
  c.nx = e;
  const uint64_t len = e - i;
  uint32_t cp = PT_INVALID;
  if (b0 >= 0xC2 && b0 <= 0xDF && len == 2) {
    cp = ((b0 & 0x1F) << 6) | (text[i + 1] & 0x3F);
  } else if ((b0 & 0xF0) == 0xE0 && len == 3) {
    cp = ((b0 & 0x0F) << 12) | ((uint32_t)(text[i + 1] & 0x3F) << 6) | (text[i + 2] & 0x3F);
    if (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF)) cp = PT_INVALID;
  } else if (b0 >= 0xF0 && b0 <= 0xF4 && len == 4) {
    cp = ((b0 & 0x07) << 18) | ((uint32_t)(text[i + 1] & 0x3F) << 12) | ((uint32_t)(text[i + 2] & 0x3F) << 6) | (text[i + 3] & 0x3F);
    if (cp < 0x10000 || cp > 0x10FFFF) cp = PT_INVALID;
  }
  c.cp = cp;
  c.cls = cp == PT_INVALID ? PT_O : pt_class(tab, cp);
  return c;
}

SWB_HD PtChar pt_none() { PtChar c; c.cp = PT_INVALID; c.cls = PT_NONE; c.at = c.nx = 0; return c; }

template <class T>
SWB_HD PtChar pt_next(const T &text, uint64_t n, const uint8_t *__restrict__ tab, const PtChar &c) {
  return c.nx < n ? pt_at(text, n, tab, c.nx) : pt_none();
}

template <class T>
SWB_HD PtChar pt_prev(const T &text, uint64_t n, const uint8_t *__restrict__ tab, const PtChar &c) {
  if (c.at == 0) return pt_none();
  uint64_t i = c.at - 1;
  while (i > 0 && pt_is_cont(text[i])) --i;
  return pt_at(text, n, tab, i);
}

// alternative (1) at the apostrophe text[a]: true iff a match starts there and the contraction letters follow; *end = index
// of the first byte behind it
template <class T>
SWB_HD bool pt_contraction(const T &text, uint64_t n, const uint8_t *__restrict__ tab, uint64_t a, uint64_t *end) {
  PtChar ap;
  ap.cp = '\''; ap.cls = PT_O; ap.at = a; ap.nx = a + 1;
  if (a + 1 < n && pt_is_cont(text[a + 1])) return false;  // malformed: not an apostrophe character
  const PtChar p = pt_prev(text, n, tab, ap);
  if (p.cls == PT_O || p.cp == ' ') return false;
  const PtChar c1 = pt_next(text, n, tab, ap);
  if (c1.cls != PT_L) return false;
  if (SWB_FOLD_S(c1.cp) || SWB_FOLD_D(c1.cp) || SWB_FOLD_M(c1.cp) || SWB_FOLD_T(c1.cp)) { *end = c1.nx; return true; }
  const PtChar c2 = pt_next(text, n, tab, c1);
  if (c2.cls != PT_L) return false;
  if ((SWB_FOLD_L(c1.cp) && SWB_FOLD_L(c2.cp)) || (SWB_FOLD_V(c1.cp) && SWB_FOLD_E(c2.cp)) || (SWB_FOLD_R(c1.cp) && SWB_FOLD_E(c2.cp))) {
    *end = c2.nx;
    return true;
  }
  return false;
}

// does a piece end behind character c?
template <class T>
SWB_HD bool pt_piece_ends(const T &text, uint64_t n, const uint8_t *__restrict__ tab, const PtChar &c) {
  const PtChar f = pt_next(text, n, tab, c);
  if (f.cls == PT_NONE) return true;
  switch (c.cls) {
    case PT_L: {
      if (f.cls != PT_L) return true;
      uint64_t end = 0;
      if (c.at >= 1 && text[c.at - 1] == '\'')  // 's 'd 'm 't
        return pt_contraction(text, n, tab, c.at - 1, &end) && end == c.nx;
      if (c.at >= 2) {                          // 'll 've 're
        const PtChar p = pt_prev(text, n, tab, c);
        if (p.cls == PT_L && p.at >= 1 && text[p.at - 1] == '\'') return pt_contraction(text, n, tab, p.at - 1, &end) && end == c.nx;
      }
      return false;
    }
    case PT_N: {
      if (f.cls != PT_N) return true;
      uint32_t before = 0;  // digits of this run in front of c
      for (PtChar p = pt_prev(text, n, tab, c); p.cls == PT_N; p = pt_prev(text, n, tab, p)) ++before;
      return (before + 1) % 3 == 0;
    }
    case PT_O: {
      if (f.cls == PT_O) return false;
      if (f.cls == PT_N) return true;
      if (f.cls == PT_S) return !pt_is_nl(f);
      const PtChar p = pt_prev(text, n, tab, c);  // f is a letter
      return p.cls == PT_O || p.cp == ' ';
    }
    default: break;
  }
  // c is whitespace
  const bool c_nl = pt_is_nl(c);
  if (c_nl) {  // CR / LF directly behind an O character (through CR / LF only) belong to that character's piece
    PtChar p = pt_prev(text, n, tab, c);
    while (p.cls == PT_S && pt_is_nl(p)) p = pt_prev(text, n, tab, p);
    if (p.cls == PT_O) return !(f.cls == PT_S && pt_is_nl(f));
  }
  uint32_t behind = 0;  // whitespace characters of the run behind c (counted up to 2)
  PtChar g = f;
  while (g.cls == PT_S) {
    if (pt_is_nl(g)) return false;  // not the last CR / LF of the run
    if (behind < 2) ++behind;
    g = pt_next(text, n, tab, g);
  }
  if (c_nl) return true;
  if (g.cls == PT_NONE) return behind == 0;  // the run reaches the end of the text: one piece (6)
  if (behind >= 2) return false;
  if (behind == 1) return true;
  return !(g.cls == PT_L) && !(c.cp == ' ' && g.cls == PT_O);
}

// what input byte i contributes to the pre-tokenised stream: itself (the trainer's four delimiter bytes mapped to 0x1C-0x1F),
// followed by one ' ' if it is the last byte of a character that ends a piece. Returns the byte count (1 or 2).
template <class T>
SWB_HD uint32_t pt_emit(const T &text, uint64_t n, const uint8_t *__restrict__ tab, uint64_t i, uint8_t *b0) {
  const uint8_t b = text[i];
  *b0 = b == ' ' ? 0x1C : b == '\t' ? 0x1D : b == '\n' ? 0x1E : b == '\r' ? 0x1F : b;
  if (i + 1 < n && pt_is_cont(text[i + 1])) return 1;  // not the last byte of its character
  uint64_t s = i;
  while (s > 0 && pt_is_cont(text[s])) --s;
  const PtChar c = pt_at(text, n, tab, s);
  return pt_piece_ends(text, n, tab, c) ? 2u : 1u;
}

// The cases that settle most bytes of ordinary text without decoding anything (by class of c and of the character f behind
// it, see the top of this file); *ends = a piece ends behind byte i. Returns false when the general rule (pt_emit) is needed.
//   * not the last byte of its character                       -> no
//   * ASCII letter, f an ASCII non-letter (or end of text)      -> yes ("c in L: f not in L")
//   * ASCII letter, f an ASCII letter, no apostrophe 1 or 2 bytes before (so c cannot end a contraction) -> no
//   * ASCII space / \t / \v / \f, f an ASCII letter             -> no (it is the prefix character of alternative (2))
template <class T>
SWB_HD bool pt_fast(const T &text, uint64_t n, uint64_t i, uint32_t *ends) {
  const uint32_t b = tx_near(text, i);
  const bool has_next = i + 1 < n;
  const uint32_t nb = has_next ? tx_near(text, i + 1) : 0u;
  if (has_next && pt_is_cont((uint8_t)nb)) { *ends = 0; return true; }
  if (b >= 0x80u) return false;
  if (!has_next) { *ends = 1; return true; }
  if (nb >= 0x80u) return false;
  // (f counts as a letter only if it is well formed: an ASCII byte with continuation bytes glued to it is a class-O character)
  const bool next_letter = ((nb | 32u) - 'a') < 26u && !(i + 2 < n && pt_is_cont(tx_near(text, i + 2)));
  if (((nb | 32u) - 'a') < 26u && !next_letter) return false;
  if (((b | 32u) - 'a') < 26u) {
    if (!next_letter) { *ends = 1; return true; }
    if ((i >= 1 && tx_near(text, i - 1) == '\'') || (i >= 2 && tx_near(text, i - 2) == '\'')) return false;
    *ends = 0;
    return true;
  }
  if (next_letter && (b == ' ' || b == '\t' || b == 0x0Bu || b == 0x0Cu)) { *ends = 0; return true; }
  return false;
}

}  // namespace swb
