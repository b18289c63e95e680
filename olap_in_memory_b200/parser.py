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
``cuda_source()`` which lowers the tree to one double-precision CUDA expression
for the fused elementwise kernel (JIT-compiled by the native library)."""
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


# name -> (arity or None for n-ary, python impl, CUDA template)
FUNCS = {
    "abs": (1, _guard(abs), "fabs({0})"),
    "ceil": (1, _guard(lambda x: x if x != x or math.isinf(x) else math.ceil(x)), "ceil({0})"),
    "floor": (1, _guard(lambda x: x if x != x or math.isinf(x) else math.floor(x)), "floor({0})"),
    "round": (1, _js_round, "floor({0} + 0.5)"),
    "trunc": (1, _guard(lambda x: x if x != x or math.isinf(x) else math.trunc(x)), "trunc({0})"),
    "sqrt": (1, _guard(lambda x: math.nan if x < 0 else math.sqrt(x)), "sqrt({0})"),
    "cbrt": (1, _guard(lambda x: math.copysign(abs(x) ** (1.0 / 3.0), x)), "cbrt({0})"),
    "exp": (1, _guard(math.exp), "exp({0})"),
    "expm1": (1, _guard(math.expm1), "expm1({0})"),
    "ln": (1, _log, "log({0})"),
    "log": (1, _log, "log({0})"),
    "log1p": (1, _guard(lambda x: -math.inf if x == -1 else math.log1p(x)), "log1p({0})"),
    "log2": (1, lambda x: _log(x) / math.log(2) if x == x and x > 0 and not math.isinf(x) else _log(x), "log2({0})"),
    "log10": (1, lambda x: math.log10(x) if x == x and x > 0 and not math.isinf(x) else _log(x), "log10({0})"),
    "lg": (1, lambda x: math.log10(x) if x == x and x > 0 and not math.isinf(x) else _log(x), "log10({0})"),
    "sin": (1, _guard(math.sin), "sin({0})"),
    "cos": (1, _guard(math.cos), "cos({0})"),
    "tan": (1, _guard(math.tan), "tan({0})"),
    "asin": (1, _guard(math.asin), "asin({0})"),
    "acos": (1, _guard(math.acos), "acos({0})"),
    "atan": (1, _guard(math.atan), "atan({0})"),
    "sinh": (1, _guard(math.sinh), "sinh({0})"),
    "cosh": (1, _guard(math.cosh), "cosh({0})"),
    "tanh": (1, _guard(math.tanh), "tanh({0})"),
    "asinh": (1, _guard(math.asinh), "asinh({0})"),
    "acosh": (1, _guard(math.acosh), "acosh({0})"),
    "atanh": (1, _guard(lambda x: math.copysign(math.inf, x) if abs(x) == 1 else math.atanh(x)), "atanh({0})"),
    "sign": (1, _sign, "olap_sign({0})"),
    "isNaN": (1, lambda x: 1.0 if x != x else 0.0, "(isnan({0}) ? 1.0 : 0.0)"),
    "pow": (2, _js_pow, "olap_pow({0}, {1})"),
    "atan2": (2, _guard(math.atan2), "atan2({0}, {1})"),
    "roundTo": (2, lambda x, n: _js_round(x * 10 ** n) / 10 ** n, "(floor({0} * pow(10.0, {1}) + 0.5) / pow(10.0, {1}))"),
    "if": (3, lambda c, a, b: a if (c == c and c != 0) else b, "(olap_truthy({0}) ? ({1}) : ({2}))"),
    "min": (None, _js_min, None),
    "max": (None, _js_max, None),
    "hypot": (None, _guard(lambda *a: math.hypot(*a)), None),
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
_BIN_CU = {
    "+": "({0} + {1})",
    "-": "({0} - {1})",
    "*": "({0} * {1})",
    "/": "({0} / {1})",
    "%": "fmod({0}, {1})",
    "^": "olap_pow({0}, {1})",
    "||": "olap_coalesce_add({0}, {1})",
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
    if v == int(v) and abs(v) < 1e15:
        return str(int(v))
    return repr(v)


def _c_literal(v):
    if v != v:
        return "olap_nan()"
    if math.isinf(v):
        return "olap_inf()" if v > 0 else "(-olap_inf())"
    return repr(float(v))


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
    def cuda_source(self, slot_of):
        """One CUDA double expression.  `slot_of` maps a variable name to the
        C identifier holding its value (a stored-measure cell or a total)."""

        def cu(n):
            tag = n[0]
            if tag == "num":
                return _c_literal(n[1])
            if tag == "var":
                return slot_of[n[1]]
            if tag == "neg":
                return f"(-{cu(n[1])})"
            if tag == "bin":
                return _BIN_CU[n[1]].format(cu(n[2]), cu(n[3]))
            if tag == "cond":
                return f"(olap_truthy({cu(n[1])}) ? ({cu(n[2])}) : ({cu(n[3])}))"
            if tag == "call":
                name, args = n[1], [cu(a) for a in n[2]]
                if name in ("min", "max"):
                    out = args[0] if len(args) > 1 else f"olap_{name}({args[0]}, {args[0]})"
                    for a in args[1:]:
                        out = f"olap_{name}({out}, {a})"
                    return out
                if name == "hypot":
                    return "sqrt(" + " + ".join(f"({a}) * ({a})" for a in args) + ")"
                return FUNCS[name][2].format(*args)
            raise AssertionError(tag)

        return cu(self.node)


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
