"""A small reader for the EDN / Clojure-literal subset MJPdes model files use.

The reference CLI reads ONE form from the input file with ``clojure.edn/read`` and
``eval``s it in ns ``gov.nist.MJPdes.core`` (main.clj:12-14); model files are
``(map->Model {... (map->ExpoMachine {...}) ...})`` forms (resources/example.clj).
This reader produces plain Python data:

    map      -> dict (insertion-ordered, like array-map)     vector -> list
    list     -> ``Form`` (a tuple subclass)                    set    -> frozenset (of hashables) or ``SetForm``
    keyword  -> ``Keyword`` (str subclass, includes the colon)  symbol -> ``Symbol`` (str subclass)
    ##Inf / ##-Inf / ##NaN, integers, floats, ratios (-> float), strings, true/false/nil

``#_`` discards, ``;`` comments, ``^meta`` (skipped), ``'``, ``@`` and ``#(...)`` are
tolerated so that the reference's test sources can be scanned for golden vectors.
"""
from __future__ import annotations

import math
import re


class Keyword(str):
    __slots__ = ()

    def __repr__(self):
        return str(self)

    @property
    def name(self) -> str:
        return self[1:]


class Symbol(str):
    __slots__ = ()

    def __repr__(self):
        return str(self)


class Form(tuple):
    """A Clojure list ``( ... )``."""
    __slots__ = ()

    @property
    def head(self):
        return self[0] if self else None


class SetForm(tuple):
    """A set literal whose elements are not all hashable (order of appearance)."""
    __slots__ = ()


class EdnError(ValueError):
    pass


_DELIMS = set("()[]{}\"; \t\r\n,")
_INT = re.compile(r"^[+-]?\d+N?$")
_FLOAT = re.compile(r"^[+-]?(\d+\.\d*|\d*\.\d+|\d+)([eE][+-]?\d+)?M?$")
_RATIO = re.compile(r"^([+-]?\d+)/(\d+)$")
_EOF = object()


class _Reader:
    def __init__(self, text: str):
        self.s, self.i, self.n = text, 0, len(text)

    def _skip_ws(self):
        s, n = self.s, self.n
        while self.i < n:
            c = s[self.i]
            if c in " \t\r\n,":
                self.i += 1
            elif c == ";":
                while self.i < n and s[self.i] != "\n":
                    self.i += 1
            else:
                break

    def read(self, eof_ok: bool = False):
        while True:
            self._skip_ws()
            if self.i >= self.n:
                if eof_ok:
                    return _EOF
                raise EdnError("unexpected end of input")
            c = self.s[self.i]
            if c == "(":
                self.i += 1
                return Form(self._seq(")"))
            if c == "[":
                self.i += 1
                return self._seq("]")
            if c == "{":
                self.i += 1
                items = self._seq("}")
                if len(items) % 2:
                    raise EdnError("map literal with an odd number of forms")
                return {_hashable(items[k]): items[k + 1] for k in range(0, len(items), 2)}
            if c in ")]}":
                raise EdnError(f"unbalanced {c!r} at offset {self.i}")
            if c == '"':
                return self._string()
            if c == "\\":
                return self._char()
            if c == "'":
                self.i += 1
                return Form((Symbol("quote"), self.read()))
            if c == "`":
                self.i += 1
                return Form((Symbol("syntax-quote"), self.read()))
            if c == "~":
                self.i += 2 if self.s.startswith("~@", self.i) else 1
                return Form((Symbol("unquote"), self.read()))
            if c == "@":
                self.i += 1
                return Form((Symbol("deref"), self.read()))
            if c == "^":
                self.i += 1
                self.read()          # metadata: dropped
                continue
            if c == "#":
                nxt = self.s[self.i + 1] if self.i + 1 < self.n else ""
                if nxt == "_":
                    self.i += 2
                    self.read()      # discard
                    continue
                if nxt == "{":
                    self.i += 2
                    items = self._seq("}")
                    if any(isinstance(x, (dict, list)) for x in items):
                        return SetForm(items)
                    return frozenset(items)
                if nxt == "(":
                    self.i += 2
                    return Form((Symbol("fn*"),) + tuple(self._seq(")")))
                if nxt == '"':
                    self.i += 1
                    return self._string()
                if nxt == "#":
                    self.i += 2
                    tok = self._token()
                    if tok == "Inf":
                        return math.inf
                    if tok == "-Inf":
                        return -math.inf
                    if tok == "NaN":
                        return math.nan
                    raise EdnError(f"unknown symbolic value ##{tok}")
                if nxt == "'":
                    self.i += 2
                    return Form((Symbol("var"), self.read()))
                # tagged literal: #tag form -> (tag form)
                self.i += 1
                tag = self._token()
                return Form((Symbol("#" + tag), self.read()))
            return self._atom(self._token())

    def _seq(self, close: str) -> list:
        out = []
        while True:
            self._skip_ws()
            if self.i >= self.n:
                raise EdnError(f"missing {close!r}")
            if self.s[self.i] == close:
                self.i += 1
                return out
            out.append(self.read())

    def _token(self) -> str:
        j = self.i
        while j < self.n and self.s[j] not in _DELIMS:
            j += 1
        tok = self.s[self.i:j]
        if not tok:
            raise EdnError(f"empty token at offset {self.i}")
        self.i = j
        return tok

    def _string(self) -> str:
        assert self.s[self.i] == '"'
        j = self.i + 1
        out = []
        esc = {"n": "\n", "t": "\t", "r": "\r", '"': '"', "\\": "\\"}
        while j < self.n and self.s[j] != '"':
            if self.s[j] == "\\":
                j += 1
                out.append(esc.get(self.s[j], self.s[j]))
            else:
                out.append(self.s[j])
            j += 1
        if j >= self.n:
            raise EdnError("unterminated string")
        self.i = j + 1
        return "".join(out)

    def _char(self) -> str:
        self.i += 1
        j = self.i + 1
        while j < self.n and self.s[j] not in _DELIMS:
            j += 1
        tok = self.s[self.i:j]
        self.i = j
        return {"newline": "\n", "space": " ", "tab": "\t"}.get(tok, tok[:1])

    @staticmethod
    def _atom(tok: str):
        if tok == "nil":
            return None
        if tok == "true":
            return True
        if tok == "false":
            return False
        if tok[0] == ":":
            return Keyword(tok)
        if _INT.match(tok):
            return int(tok.rstrip("N"))
        if _FLOAT.match(tok):
            return float(tok.rstrip("M"))
        m = _RATIO.match(tok)
        if m:
            return int(m.group(1)) / int(m.group(2))
        return Symbol(tok)


def _hashable(x):
    if isinstance(x, list):
        return tuple(_hashable(y) for y in x)
    if isinstance(x, dict):
        return tuple((k, _hashable(v)) for k, v in x.items())
    return x


def read_string(text: str):
    """Read the first form of ``text`` (what ``clojure.edn/read`` does, main.clj:12)."""
    return _Reader(text).read()


def read_all(text: str) -> list:
    """Read every top-level form of ``text``."""
    r = _Reader(text)
    out = []
    while True:
        f = r.read(eof_ok=True)
        if f is _EOF:
            return out
        out.append(f)


def read_file(path: str):
    with open(path) as fh:
        return read_string(fh.read())
