"""Iteration order of Clojure maps, restated: what ``(vals m)`` / ``(map f m)`` / ``group-by`` yield once a map has
more than 8 entries.

The reference walks plain maps in three places that matter for results: ``(group-by :m ...)`` in clean-log-buf
(util/log.clj:44), ``(:jobmix model)`` in new-job (core.clj:332) and ``(vals (:blocked res))`` in calc-bneck
(core.clj:641-642).  Up to 8 entries a Clojure map is a PersistentArrayMap and iterates in insertion order; the
9th key promotes it to a PersistentHashMap (HASHTABLE_THRESHOLD = 16 array slots), a 32-way hash trie whose
nodes (BitmapIndexedNode and ArrayNode alike) enumerate their children by ascending 5-bit chunk of the key's
hash, lowest chunk first.  So the iteration order of a hash map is: keys sorted by
(h & 31, (h >>> 5) & 31, (h >>> 10) & 31, ...), independent of insertion order (keys with equal 32-bit hashes
aside, which does not happen for the short names used here).

Keyword hash, clojure 1.9 (the version pinned by the reference's project.clj):
    Keyword.hasheq = Symbol.hasheq + 0x9e3779b9
    Symbol.hasheq  = Util.hashCombine(Murmur3.hashUnencodedChars(name), Util.hash(ns))      ; ns nil -> 0
    hashCombine(seed, h) = seed ^ (h + 0x9e3779b9 + (seed << 6) + (seed >> 2))              ; int arithmetic
Restated from the published Clojure sources (the JVM is not available here to execute them); the known value
(hash :a) = -2123407586 is reproduced (tests/test_cljhash.py).
"""
from __future__ import annotations

M32 = 0xFFFFFFFF


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (32 - r))) & M32


def _mix_k1(k1: int) -> int:
    k1 = (k1 * 0xCC9E2D51) & M32
    k1 = _rotl(k1, 15)
    return (k1 * 0x1B873593) & M32


def _mix_h1(h1: int, k1: int) -> int:
    h1 ^= k1
    h1 = _rotl(h1, 13)
    return (h1 * 5 + 0xE6546B64) & M32


def _fmix(h1: int, length: int) -> int:
    h1 ^= length
    h1 ^= h1 >> 16
    h1 = (h1 * 0x85EBCA6B) & M32
    h1 ^= h1 >> 13
    h1 = (h1 * 0xC2B2AE35) & M32
    h1 ^= h1 >> 16
    return h1


def murmur3_unencoded_chars(s: str) -> int:
    """clojure.lang.Murmur3.hashUnencodedChars: UTF-16 code units two at a time, seed 0."""
    u = s.encode("utf-16-le")
    units = [u[i] | (u[i + 1] << 8) for i in range(0, len(u), 2)]
    h1 = 0
    for i in range(1, len(units), 2):
        h1 = _mix_h1(h1, _mix_k1(units[i - 1] | (units[i] << 16)))
    if len(units) & 1:
        h1 ^= _mix_k1(units[-1])
    return _fmix(h1, 2 * len(units))


def _sar(x: int, r: int) -> int:                     # arithmetic shift of a 32-bit int
    return ((x - (1 << 32)) >> r) & M32 if x & 0x80000000 else x >> r


def _hash_combine(seed: int, h: int) -> int:
    return (seed ^ ((h + 0x9E3779B9 + ((seed << 6) & M32) + _sar(seed, 2)) & M32)) & M32


def java_string_hash(s: str) -> int:
    """String.hashCode over UTF-16 code units: what Util.hash(Object) returns for the namespace string."""
    u = s.encode("utf-16-le")
    h = 0
    for i in range(0, len(u), 2):
        h = (31 * h + (u[i] | (u[i + 1] << 8))) & M32
    return h


def symbol_hash(name: str, ns: str | None = None) -> int:
    return _hash_combine(murmur3_unencoded_chars(name), 0 if ns is None else java_string_hash(ns))


def keyword_hash(kw: str) -> int:
    """hasheq of a keyword given as ':name' or ':ns/name' (unsigned 32-bit)."""
    body = kw[1:] if kw.startswith(":") else kw
    ns, name = (body.split("/", 1) if "/" in body and body != "/" else (None, body))
    return (symbol_hash(name, ns) + 0x9E3779B9) & M32


def to_signed(h: int) -> int:
    return h - (1 << 32) if h & 0x80000000 else h


def trie_key(h: int) -> tuple:
    """Sort key of a hash in PersistentHashMap iteration order."""
    return tuple((h >> s) & 31 for s in range(0, 35, 5))


def map_order(keys: list) -> list:
    """The order in which a Clojure map holding `keys` (inserted in the given order) iterates."""
    keys = list(keys)
    if len(keys) <= 8:
        return keys
    return sorted(keys, key=lambda k: trie_key(keyword_hash(k)))


def hash_rank(keys: list) -> list:
    """rank[i] = position of keys[i] among `keys` in HASH-MAP order (whatever their number)."""
    order = sorted(range(len(keys)), key=lambda i: trie_key(keyword_hash(keys[i])))
    rank = [0] * len(keys)
    for pos, i in enumerate(order):
        rank[i] = pos
    return rank
