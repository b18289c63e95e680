"""Formula parser for computed measures (host side).

The reference builds an ``expr-eval`` Parser with logical/comparison/in/
assignment switched off, adds ``isNaN`` and re-purposes ``||`` as a
NaN-coalescing add (/root/reference/src/parser.js:3-26).  ``@growblocks/
expr-eval`` (git tag v2.0.6) is not vendored in the reference, so this module
restates the arithmetic subset of its published grammar:

    ternary  c ? a : b            (lowest)
    additive + - ||               (left associative; || is parser.js:18-23)
    term     * / %
    unary    - +  and prefix functions (``sqrt x``)
    power    ^                    (right associative, binds tighter than unary -)
    call     f(a, b, ...), atoms: number, variable, PI, E, ( ... )

Only ``+``, ``/`` and ``||`` are pinned by the reference's own tests
(test/cube-accessors.js:58-67, test/cube-to-cube.js:360-382); everything else
is parity-unpinned and follows JavaScript ``Math`` semantics.

An Expression offers the members the cube uses (src/cube.js:123-124, 253-256,
336-338, 359, 1145): variables(), evaluate(), toString(), substitute(); plus
``postfix()``, the serialised instruction list handed to olap_eval(), which the
native library lowers to one fused double-precision sm_100a kernel (NVRTC)."""
from __future__ import annotations

import math
import re

_TOKEN = re.compile(
    r"\s*(?:(?P<num>(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)|(?P<id>[A-Za-z_][A-Za-z0-9_]*)|(?P<op>\|\||[-+*/%^(),?:]))"
)

CONSTS = {"PI": math.pi, "E": math.e}


def _js_round(x):
    if x != x or math.isinf(x):
        return x
    return float(math.floor(x + 0.5))


def _js_max(*a):
    r = -math.inf
    for v in a:
        if v != v:
            return math.nan
        if v > r or (v == 0 and r == 0 and math.copysign(1, v) > 0):
            r = v
    return r


def _js_min(*a):
    r = math.inf
    for v in a:
        if v != v:
            return math.nan
        if v < r or (v == 0 and r == 0 and math.copysign(1, v) < 0):
            r = v
    return r


def _js_div(a, b):
    if b == 0:
        if a != a or a == 0:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1, b)
    return a / b


def _js_mod(a, b):
    if b == 0 or a != a or b != b or math.isinf(a):
        return math.nan
    if math.isinf(b):
        return a
    return math.fmod(a, b)


def _js_pow(a, b):
    try:
        if b != b:
            return math.nan
        if b == 0:
            return 1.0
        if a != a:
            return math.nan
        # JS: (+-1) ** +-Infinity is NaN (C pow gives 1)
        if abs(a) == 1 and math.isinf(b):
            return math.nan
        return math.pow(a, b)
    except OverflowError:
        return math.inf if a > 0 or float(b).is_integer() and int(b) % 2 == 0 else -math.inf
    except ValueError:
        return math.nan
    except ZeroDivisionError:
        return math.inf


def _guard(fn):
    def wrapped(*a):
        try:
            return float(fn(*a))
        except (ValueError, ZeroDivisionError):
            return math.nan
        except OverflowError:
            return math.inf

    return wrapped


def _coalesce_add(a, b):
    """parser.js:18-23."""
    if a != a and b == b:
        return b
    if a == a and b != b:
        return a
    return a + b


def _sign(x):
    if x != x:
        return math.nan
    return (x > 0) - (x < 0) or x


def _log(x):
    if x != x:
        return math.nan
    if x == 0:
        return -math.inf
    if x < 0:
        return math.nan
    return math.log(x) if not math.isinf(x) else math.inf


# name -> (arity or None for n-ary, python impl with JS Math semantics)
FUNCS = {
    "abs": (1, _guard(abs)),
    "ceil": (1, _guard(lambda x: x if x != x or math.isinf(x) else math.ceil(x))),
    "floor": (1, _guard(lambda x: x if x != x or math.isinf(x) else math.floor(x))),
    "round": (1, _js_round),
    "trunc": (1, _guard(lambda x: x if x != x or math.isinf(x) else math.trunc(x))),
    "sqrt": (1, _guard(lambda x: math.nan if x < 0 else math.sqrt(x))),
    "cbrt": (1, _guard(lambda x: math.copysign(abs(x) ** (1.0 / 3.0), x))),
    "exp": (1, _guard(math.exp)),
    "expm1": (1, _guard(math.expm1)),
    "ln": (1, _log),
    "log": (1, _log),
    "log1p": (1, _guard(lambda x: -math.inf if x == -1 else math.log1p(x))),
    "log2": (1, lambda x: _log(x) / math.log(2) if x == x and x > 0 and not math.isinf(x) else _log(x)),
    "log10": (1, lambda x: math.log10(x) if x == x and x > 0 and not math.isinf(x) else _log(x)),
    "lg": (1, lambda x: math.log10(x) if x == x and x > 0 and not math.isinf(x) else _log(x)),
    "sin": (1, _guard(math.sin)),
    "cos": (1, _guard(math.cos)),
    "tan": (1, _guard(math.tan)),
    "asin": (1, _guard(math.asin)),
    "acos": (1, _guard(math.acos)),
    "atan": (1, _guard(math.atan)),
    "sinh": (1, _guard(math.sinh)),
    "cosh": (1, _guard(math.cosh)),
    "tanh": (1, _guard(math.tanh)),
    "asinh": (1, _guard(math.asinh)),
    "acosh": (1, _guard(math.acosh)),
    "atanh": (1, _guard(lambda x: math.copysign(math.inf, x) if abs(x) == 1 else math.atanh(x))),
    "sign": (1, _sign),
    "isNaN": (1, lambda x: 1.0 if x != x else 0.0),
    "pow": (2, _js_pow),
    "atan2": (2, _guard(math.atan2)),
    "roundTo": (2, lambda x, n: _js_round(x * 10 ** n) / 10 ** n),
    "if": (3, lambda c, a, b: a if (c == c and c != 0) else b),
    "min": (None, _js_min),
    "max": (None, _js_max),
    "hypot": (None, _guard(lambda *a: math.hypot(*a))),
}

PREFIX_FUNCS = {k for k, v in FUNCS.items() if v[0] == 1 and k != "isNaN"}

_BIN_PY = {
    "+": lambda a, b: a + b,
    "-": lambda a, b: a - b,
    "*": lambda a, b: a * b,
    "/": _js_div,
    "%": _js_mod,
    "^": _js_pow,
    "||": _coalesce_add,
}

class ParseError(ValueError):
    pass


# AST nodes are tuples: ("num", v) ("var", name) ("neg", a) ("bin", op, a, b)
# ("call", name, [args]) ("cond", c, a, b)
class _P:
    def __init__(self, text):
        self.toks = []
        pos = 0
        text = text.strip()
        while pos < len(text):
            m = _TOKEN.match(text, pos)
            if not m or m.end() == pos:
                raise ParseError(f"parse error [1:{pos + 1}]: Unknown character \"{text[pos]}\"")
            pos = m.end()
            kind = m.lastgroup
            self.toks.append((kind, m.group(kind)))
        self.i = 0

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else (None, None)

    def take(self, op=None):
        kind, val = self.peek()
        if op is not None and not (kind == "op" and val == op):
            raise ParseError(f"parse error: Expected {op}")
        self.i += 1
        return kind, val

    def is_op(self, *ops):
        kind, val = self.peek()
        return kind == "op" and val in ops

    def expression(self):
        node = self.conditional()
        return node

    def conditional(self):
        cond = self.additive()
        if self.is_op("?"):
            self.take()
            yes = self.conditional()
            self.take(":")
            no = self.conditional()
            return ("cond", cond, yes, no)
        return cond

    def additive(self):
        left = self.term()
        while self.is_op("+", "-", "||"):
            _, op = self.take()
            left = ("bin", op, left, self.term())
        return left

    def term(self):
        left = self.factor()
        while self.is_op("*", "/", "%"):
            _, op = self.take()
            left = ("bin", op, left, self.factor())
        return left

    def factor(self):
        if self.is_op("-"):
            self.take()
            return ("neg", self.factor())
        if self.is_op("+"):
            self.take()
            return self.factor()
        kind, val = self.peek()
        if kind == "id" and val in PREFIX_FUNCS:
            nxt = self.toks[self.i + 1] if self.i + 1 < len(self.toks) else (None, None)
            if nxt != ("op", "("):
                self.take()
                return ("call", val, [self.factor()])
        return self.power()

    def power(self):
        base = self.call()
        if self.is_op("^"):
            self.take()
            return ("bin", "^", base, self.factor())
        return base

    def call(self):
        kind, val = self.peek()
        if kind == "id" and self.i + 1 < len(self.toks) and self.toks[self.i + 1] == ("op", "("):
            if val not in FUNCS:
                raise ParseError(f"parse error: unknown function {val}")
            self.take()
            self.take("(")
            args = []
            if not self.is_op(")"):
                args.append(self.expression())
                while self.is_op(","):
                    self.take()
                    args.append(self.expression())
            self.take(")")
            arity = FUNCS[val][0]
            if arity is not None and arity != len(args):
                raise ParseError(f"parse error: {val} expects {arity} argument(s)")
            if arity is None and not args:
                raise ParseError(f"parse error: {val} expects at least one argument")
            return ("call", val, args)
        return self.atom()

    def atom(self):
        kind, val = self.peek()
        if kind == "num":
            self.take()
            return ("num", float(val))
        if kind == "id":
            self.take()
            if val in CONSTS:
                return ("num", CONSTS[val])
            return ("var", val)
        if kind == "op" and val == "(":
            self.take()
            node = self.expression()
            self.take(")")
            return node
        raise ParseError(f"parse error: unexpected {val!r}")


def _fmt_num(v):
    if v != v:
        return "NaN"
    if v in (math.inf, -math.inf):  # a literal such as 1e999
        return "Infinity" if v > 0 else "-Infinity"
    if v == int(v) and abs(v) < 1e15:
        return str(int(v))
    return repr(v)


class Expression:
    def __init__(self, node):
        self.node = node

    # --- reference Expression API ------------------------------------
    def variables(self, _options=None, **_kw):
        seen = []

        def walk(n):
            tag = n[0]
            if tag == "var":
                if n[1] not in seen:
                    seen.append(n[1])
            elif tag == "neg":
                walk(n[1])
            elif tag == "bin":
                walk(n[2]), walk(n[3])
            elif tag == "call":
                for a in n[2]:
                    walk(a)
            elif tag == "cond":
                walk(n[1]), walk(n[2]), walk(n[3])

        walk(self.node)
        return seen

    def evaluate(self, params=None):
        params = params or {}

        def ev(n):
            tag = n[0]
            if tag == "num":
                return n[1]
            if tag == "var":
                if n[1] not in params:
                    raise KeyError(f"undefined variable: {n[1]}")
                return float(params[n[1]])
            if tag == "neg":
                return -ev(n[1])
            if tag == "bin":
                return _BIN_PY[n[1]](ev(n[2]), ev(n[3]))
            if tag == "call":
                return FUNCS[n[1]][1](*[ev(a) for a in n[2]])
            if tag == "cond":
                c = ev(n[1])
                return ev(n[2]) if (c == c and c != 0) else ev(n[3])
            raise AssertionError(tag)

        return ev(self.node)

    def toString(self):
        def s(n):
            tag = n[0]
            if tag == "num":
                return _fmt_num(n[1]) if n[1] >= 0 else f"({_fmt_num(n[1])})"
            if tag == "var":
                return n[1]
            if tag == "neg":
                return f"(-{s(n[1])})"
            if tag == "bin":
                return f"({s(n[2])} {n[1]} {s(n[3])})"
            if tag == "call":
                return f"{n[1]}({', '.join(s(a) for a in n[2])})"
            if tag == "cond":
                return f"({s(n[1])} ? ({s(n[2])}) : ({s(n[3])}))"
            raise AssertionError(tag)

        return s(self.node)

    __str__ = toString

    def substitute(self, variable, expr):
        repl = expr.node if isinstance(expr, Expression) else getParser().parse(str(expr)).node

        def sub(n):
            tag = n[0]
            if tag == "var":
                return repl if n[1] == variable else n
            if tag == "neg":
                return ("neg", sub(n[1]))
            if tag == "bin":
                return ("bin", n[1], sub(n[2]), sub(n[3]))
            if tag == "call":
                return ("call", n[1], [sub(a) for a in n[2]])
            if tag == "cond":
                return ("cond", sub(n[1]), sub(n[2]), sub(n[3]))
            return n

        return Expression(sub(self.node))

    # --- lowering for the fused device kernel -------------------------
    def postfix(self, slot_of):
        """The formula as the space separated postfix text olap_eval() takes
        (include/olap_gpu.h): `slot_of` maps a variable name to ``v<k>`` (cell of
        input store k) or ``t<k>`` (k-th total)."""
        out = []

        def emit(n):
            tag = n[0]
            if tag == "num":
                v = n[1]
                out.append("#nan" if v != v else "#inf" if v == math.inf else "#-inf" if v == -math.inf else f"#{v!r}")
            elif tag == "var":
                out.append(slot_of[n[1]])
            elif tag == "neg":
                emit(n[1])
                out.append("neg")
            elif tag == "bin":
                emit(n[2])
                emit(n[3])
                out.append(n[1])
            elif tag == "cond":
                emit(n[1])
                emit(n[2])
                emit(n[3])
                out.append("?:")
            elif tag == "call":
                for a in n[2]:
                    emit(a)
                out.append(f"call:{n[1]}:{len(n[2])}")
            else:
                raise AssertionError(tag)

        emit(self.node)
        return " ".join(out)


class Parser:
    def parse(self, text):
        p = _P(text)
        if not p.toks:
            raise ParseError("parse error: empty expression")
        node = p.expression()
        if p.i != len(p.toks):
            raise ParseError(f"parse error: unexpected {p.peek()[1]!r}")
        return Expression(node)


def getParser():
    return Parser()
