"""Boolean expression evaluator for `logic_func_data` strings.

Same language as the reference's shunting-yard evaluator (gym_PBN/utils/logic/eval.py:47-167):
operators `not` > `and` > `or` (left associative), parentheses, literals `True` / `False`, symbols
looked up in a dictionary.  Implemented here as a small recursive-descent compiler to a closure tree,
so a function is parsed once and evaluated 2^k times by the network compiler.
"""
import re

_TOKEN = re.compile(r"\(|\)|[^\s()]+")
_KEYWORDS = {"and", "or", "not", "(", ")", "True", "False"}
_SYMBOL = re.compile(r"[a-zA-Z]+\d*")


def _tokens(text):
    toks = _TOKEN.findall(text)
    for t in toks:
        if t not in _KEYWORDS and not _SYMBOL.match(t):
            raise Exception(f"Illegal token {t}")
    return toks


class _Parser:
    def __init__(self, toks):
        self.toks, self.pos = toks, 0

    def peek(self):
        return self.toks[self.pos] if self.pos < len(self.toks) else None

    def take(self):
        t = self.peek()
        self.pos += 1
        return t

    def parse(self):
        node = self.disjunction()
        if self.peek() is not None:
            raise Exception(f"Invalid syntax at {self.peek()} ({self.pos})")
        return node

    def disjunction(self):
        left = self.conjunction()
        while self.peek() == "or":
            self.take()
            right = self.conjunction()
            left = (lambda a, b: lambda env: a(env) or b(env))(left, right)
        return left

    def conjunction(self):
        left = self.negation()
        while self.peek() == "and":
            self.take()
            right = self.negation()
            left = (lambda a, b: lambda env: a(env) and b(env))(left, right)
        return left

    def negation(self):
        if self.peek() == "not":
            self.take()
            inner = self.negation()
            return lambda env: not inner(env)
        return self.atom()

    def atom(self):
        t = self.take()
        if t is None:
            raise Exception("Unexpected end of expression")
        if t == "(":
            node = self.disjunction()
            if self.take() != ")":
                raise Exception("Missing parenthesis")
            return node
        if t == "True":
            return lambda env: True
        if t == "False":
            return lambda env: False
        if t in _KEYWORDS:
            raise Exception(f"Invalid syntax at {t} ({self.pos - 1})")

        def lookup(env, name=t):
            if name not in env:
                raise Exception(f"Symbol {name} doesn't exist.")
            return env[name]

        return lookup


def compile_expression(text):
    """Parse once; returns f(dict) -> truthy value."""
    if not text:
        raise Exception("Empty expression string")
    return _Parser(_tokens(text)).parse()


class LogicExpressionEvaluator:
    """API-compatible front: `.dictionary`, `.evaluate(expr)`, `.get_symbols(expr)`."""

    def __init__(self, role_dict):
        self.dictionary = role_dict
        self._cache = {}

    @classmethod
    def get_symbols(cls, in_str):
        return [t for t in _tokens(in_str) if t not in _KEYWORDS]

    def evaluate(self, in_str):
        fn = self._cache.get(in_str)
        if fn is None:
            fn = self._cache[in_str] = compile_expression(in_str)
        return fn(self.dictionary)


if __name__ == "__main__":
    ev = LogicExpressionEvaluator({"u": False, "x1": False, "x2": False, "x3": True, "x4": False})
    print(ev.evaluate("not x4 and not u and (x2 or x3)"))
