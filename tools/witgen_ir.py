#!/usr/bin/env python3
"""Circuit IR for the rv32im witness generator.

The reference ships the rv32im circuit's per-cycle step functions only as zirgen-generated C++ / CUDA
(risc0/circuit/rv32im-sys/kernels/cxx/{steps.cpp,types.h.inc,layout.cpp.inc}: `step_Top` fills the data columns of one
cycle from the preflight trace, `step_TopAccum` the accum columns from data + mix). That generated code IS the circuit
definition (there is no other machine-readable form of it in the tree), so - exactly as tools/circuit_ir.py does for
the constraint polynomial - this tool parses it into a neutral IR (functions / statements / expressions as JSON, value
types, and the column layouts flattened to integer tables) that is committed as
risc0_b200/circuits/rv32im_witgen.ir.json.gz. tools/gen_witgen.py emits this backend's own CUDA from the IR; nothing
of the reference's text is carried over (comments, source locations and the C++ helper runtime are dropped).

    python tools/witgen_ir.py [--ref /root/reference] [-o risc0_b200/circuits/rv32im_witgen.ir.json.gz]

IR (JSON):
  types:    {name: ["struct", [[field, type], ...]] | ["array", elem_type, n]}      value AND layout types
  layouts:  {name: {"type": T, "cols": [flattened column numbers]}}                 kLayout_Top, kLayoutGlobal, ...
  funcs:    {name: {"ret": T, "params": [[T, name], ...], "body": [stmt, ...]}}     ExecContext parameter dropped
  stmt:     ["decl", T, name, expr|null] | ["unpack", [names], expr] | ["assign", name, expr]
            | ["if", [[cond, [stmt]], ...], [stmt]|null] | ["eqz", expr, site] | ["store", lay, expr]
            | ["storeext", lay, expr] | ["expr", expr] | ["return", expr|null] | ["unreachable"]
  expr:     ["val", n] | ["ext", a, b, c, d] | ["var", name] | ["fld", e, name] | ["idx", e, e] | ["bin", op, a, b]
            | ["neg", e] | ["call", fn, [e]] | ["load", lay, back] | ["loadext", lay, back] | ["ll", lay, path]
            | ["ls", lay, e] | ["bind", layout_name, buf] | ["struct", T, [[f, e]]] | ["array", T, [e]]
            | ["map", arr, lay, lambda] | ["reduce", arr, init, lay, lambda] | ["extern", name, [e]] | ["num", n]
  lambda:   {"params": [[T, name], ...], "body": [stmt]}
  path:     [["f", name] | ["i", n], ...]
"""
import argparse
import gzip
import json
import os
import re
import sys

TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<lcomment>//[^\n]*)
  | (?P<bcomment>/\*.*?\*/)
  | (?P<str>"(?:[^"\\]|\\.)*")
  | (?P<num>\d+)
  | (?P<id>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op>::|&&|\[&\]|[{}()\[\],;.=+\-*<>&])
""", re.S | re.X)


def tokenize(text):
    out = []
    pos = 0
    n = len(text)
    while pos < n:
        m = TOKEN_RE.match(text, pos)
        if not m:
            raise SyntaxError("cannot tokenize at %d: %r" % (pos, text[pos:pos + 40]))
        pos = m.end()
        k = m.lastgroup
        if k in ("ws", "lcomment", "bcomment"):
            continue
        out.append((k, m.group()))
    return out


class Parser:
    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def accept(self, val):
        if self.peek()[1] == val:
            self.i += 1
            return True
        return False

    def expect(self, val):
        tok = self.next()
        if tok[1] != val:
            ctx = " ".join(t[1] for t in self.t[max(0, self.i - 12):self.i + 6])
            raise SyntaxError("expected %r, got %r near: %s" % (val, tok[1], ctx))
        return tok

    def ident(self):
        tok = self.next()
        if tok[0] != "id":
            ctx = " ".join(t[1] for t in self.t[max(0, self.i - 12):self.i + 6])
            raise SyntaxError("expected identifier, got %r near: %s" % (tok[1], ctx))
        return tok[1]

    # ---- types: Name | Name::value_type | BoundLayout<Name[::value_type]> | std::array<T, N>
    def parse_type(self):
        name = self.ident()
        if name == "std" and self.peek()[1] == "::":
            self.next()
            sub = self.ident()
            assert sub in ("array", "initializer_list"), sub
            self.expect("<")
            inner = self.parse_type()
            if sub == "array":
                self.expect(",")
                n = int(self.next()[1])
                self.expect(">")
                return "array<%s,%d>" % (inner, n)
            self.expect(">")
            return "initlist<%s>" % inner
        if name == "BoundLayout":
            self.expect("<")
            inner = self.parse_type()
            self.expect(">")
            return "BoundLayout<%s>" % inner
        if self.peek()[1] == "::":
            self.next()
            sub = self.ident()
            assert sub == "value_type", sub
            return name + "::value_type"
        return name

    # ---- expressions
    def parse_expr(self):
        return self.parse_sum()

    def parse_sum(self):
        e = self.parse_term()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            r = self.parse_term()
            e = ["bin", op, e, r]
        return e

    def parse_term(self):
        e = self.parse_unary()
        while self.peek()[1] == "*":
            self.next()
            r = self.parse_unary()
            e = ["bin", "*", e, r]
        return e

    def parse_unary(self):
        if self.accept("-"):
            return ["neg", self.parse_unary()]
        return self.parse_postfix()

    def parse_postfix(self):
        e = self.parse_primary()
        while True:
            if self.accept("."):
                e = ["fld", e, self.ident()]
            elif self.accept("["):
                idx = self.parse_expr()
                self.expect("]")
                e = ["idx", e, idx]
            else:
                return e

    def parse_args(self):
        args = []
        if self.accept(")"):
            return args
        while True:
            args.append(self.parse_expr())
            if self.accept(")"):
                return args
            self.expect(",")

    def parse_path(self, e):
        """a.b[0]._super parsed as an expression -> [["f","a"],["f","b"],["i",0],["f","_super"]]"""
        if e[0] == "var":
            return [["f", e[1]]]
        if e[0] == "fld":
            return self.parse_path(e[1]) + [["f", e[2]]]
        if e[0] == "idx":
            assert e[2][0] == "num", e
            return self.parse_path(e[1]) + [["i", e[2][1]]]
        raise SyntaxError("bad layout path %r" % (e,))

    def parse_lambda(self):
        self.expect("[&]")
        self.expect("(")
        params = []
        while not self.accept(")"):
            ty = self.parse_type()
            params.append([ty, self.ident()])
            self.accept(",")
        body = self.parse_block()
        return {"params": params, "body": body}

    def parse_primary(self):
        kind, val = self.peek()
        if kind == "num":
            self.next()
            return ["num", int(val)]
        if kind == "str":
            self.next()
            return ["str", val[1:-1]]
        if val == "(":
            self.next()
            if self.peek()[1] == "[&]":
                lam = self.parse_lambda()
                self.expect(")")
                return ["lambda", lam]
            e = self.parse_expr()
            self.expect(")")
            return e
        if kind != "id":
            ctx = " ".join(t[1] for t in self.t[max(0, self.i - 12):self.i + 6])
            raise SyntaxError("unexpected token %r near: %s" % (val, ctx))
        # identifier-led forms
        if val == "std":
            ty = self.parse_type()
            self.expect("{")
            items = []
            while not self.accept("}"):
                items.append(self.parse_expr())
                self.accept(",")
            return ["initlist", items]
        name = self.ident()
        nxt = self.peek()[1]
        if nxt == "(":
            self.next()
            if name == "Val":
                args = self.parse_args()
                assert len(args) == 1 and args[0][0] == "num", args
                return ["val", args[0][1]]
            if name == "ExtVal":
                args = self.parse_args()
                assert len(args) == 4 and all(a[0] == "num" for a in args), args
                return ["ext"] + [a[1] for a in args]
            if name in ("LOAD", "LOAD_EXT"):
                args = self.parse_args()
                assert len(args) == 2
                return ["load" if name == "LOAD" else "loadext", args[0], args[1]]
            if name == "LAYOUT_LOOKUP":
                lay = self.parse_expr()
                self.expect(",")
                path = self.parse_path(self.parse_expr())
                self.expect(")")
                return ["ll", lay, path]
            if name == "LAYOUT_SUBSCRIPT":
                args = self.parse_args()
                assert len(args) == 2
                return ["ls", args[0], args[1]]
            if name == "BIND_LAYOUT":
                args = self.parse_args()
                assert len(args) == 2 and args[0][0] == "var" and args[1][0] == "var"
                return ["bind", args[0][1], args[1][1]]
            if name == "INVOKE_EXTERN":
                args = self.parse_args()
                assert args[0] == ["var", "ctx"] and args[1][0] == "var"
                return ["extern", args[1][1], args[2:]]
            if name == "map":
                args = self.parse_args()
                assert len(args) == 3 and args[2][0] == "lambda", args
                return ["map", args[0], args[1], args[2][1]]
            if name == "reduce":
                args = self.parse_args()
                assert len(args) == 4 and args[3][0] == "lambda"
                return ["reduce", args[0], args[1], args[2], args[3][1]]
            args = self.parse_args()
            if args and args[0] == ["var", "ctx"]:
                args = args[1:]
            return ["call", name, args]
        if nxt == "{":
            self.next()
            # struct literal (designated) or array literal (positional)
            if self.peek()[1] == ".":
                fields = []
                while not self.accept("}"):
                    self.expect(".")
                    f = self.ident()
                    self.expect("=")
                    fields.append([f, self.parse_expr()])
                    self.accept(",")
                return ["struct", name, fields]
            items = []
            while not self.accept("}"):
                items.append(self.parse_expr())
                self.accept(",")
            if not items and not name.endswith("Array"):
                return ["struct", name, []]
            return ["array", name, items]
        return ["var", name]

    # ---- statements
    def parse_block(self):
        self.expect("{")
        stmts = []
        while not self.accept("}"):
            s = self.parse_stmt()
            if s is not None:
                stmts.append(s)
        return stmts

    def skip_to_semicolon(self):
        depth = 0
        while True:
            v = self.next()[1]
            if v in "({[":
                depth += 1
            elif v in ")}]":
                depth -= 1
            elif v == ";" and depth == 0:
                return

    def parse_stmt(self):
        kind, val = self.peek()
        if val == ";":
            self.next()
            return None
        if val == "return":
            self.next()
            if self.accept(";"):
                return ["return", None]
            e = self.parse_expr()
            self.expect(";")
            return ["return", e]
        if val == "if":
            arms = []
            els = None
            while True:
                self.expect("if")
                self.expect("(")
                cond = self.parse_expr()
                self.expect(")")
                arms.append([cond, self.parse_block()])
                if self.accept("else"):
                    if self.peek()[1] == "if":
                        continue
                    els = self.parse_block()
                break
            return ["if", arms, els]
        if val == "assert":
            self.skip_to_semicolon()
            return ["unreachable"]
        if val == "EQZ":
            self.next()
            self.expect("(")
            e = self.parse_expr()
            self.expect(",")
            loc = self.next()
            assert loc[0] == "str"
            self.expect(")")
            self.expect(";")
            return ["eqz", e, loc[1][1:-1]]
        if val in ("STORE", "STORE_EXT"):
            self.next()
            self.expect("(")
            args = self.parse_args()
            self.expect(";")
            assert len(args) == 2
            return ["store" if val == "STORE" else "storeext", args[0], args[1]]
        if val == "auto":
            self.next()
            self.expect("[")
            names = []
            while not self.accept("]"):
                names.append(self.ident())
                self.accept(",")
            self.expect("=")
            e = self.parse_expr()
            self.expect(";")
            return ["unpack", names, e]
        if kind == "id":
            # declaration `T name [= e];`, assignment `name = e;`, or expression statement
            save = self.i
            if self.peek(1)[1] == "=" and self.peek(1)[0] == "op":
                name = self.ident()
                self.expect("=")
                e = self.parse_expr()
                self.expect(";")
                return ["assign", name, e]
            if val not in ("INVOKE_EXTERN",):
                try:
                    ty = self.parse_type()
                    if self.peek()[0] == "id" and self.peek(1)[1] in ("=", ";"):
                        name = self.ident()
                        if self.accept(";"):
                            return ["decl", ty, name, None]
                        self.expect("=")
                        e = self.parse_expr()
                        self.expect(";")
                        return ["decl", ty, name, e]
                except (SyntaxError, AssertionError):
                    pass
                self.i = save
        e = self.parse_expr()
        self.expect(";")
        return ["expr", e]

    # ---- top level of steps.cpp
    def parse_functions(self):
        funcs = {}
        while self.peek()[0] != "eof":
            kind, val = self.peek()
            if val in ("namespace",):
                # namespace a::b::c {
                while self.next()[1] != "{":
                    pass
                continue
            if val == "}":
                self.next()
                continue
            ret = self.parse_type()
            name = self.ident()
            self.expect("(")
            params = []
            while not self.accept(")"):
                ty = self.parse_type()
                if self.accept("&"):
                    pass
                pname = self.ident()
                if ty != "ExecContext":
                    params.append([ty, pname])
                self.accept(",")
            body = self.parse_block()
            funcs[name] = {"ret": ret, "params": params, "body": body}
        return funcs

    # ---- types.h.inc
    def parse_types(self):
        types = {}
        while self.peek()[0] != "eof":
            val = self.peek()[1]
            if val == "struct":
                self.next()
                name = self.ident()
                self.expect("{")
                fields = []
                while not self.accept("}"):
                    ty = self.parse_type()
                    fields.append([self.ident(), ty])
                    self.expect(";")
                self.expect(";")
                types[name] = ["struct", fields]
            elif val == "using":
                self.next()
                name = self.ident()
                self.expect("=")
                ty = self.parse_type()
                self.expect(";")
                m = re.match(r"array<(.*),(\d+)>$", ty)
                assert m, ty
                types[name] = ["array", m.group(1), int(m.group(2))]
            else:
                raise SyntaxError("types: unexpected %r" % val)
        return types

    # ---- layout.cpp.inc: constexpr T name = init;
    def parse_layout_init(self, consts):
        kind, val = self.peek()
        if kind == "num":
            self.next()
            return int(val)
        name = self.ident()
        if self.peek()[1] != "{":
            return consts[name]["value"]
        self.next()
        if self.peek()[1] == ".":
            fields = {}
            while not self.accept("}"):
                self.expect(".")
                f = self.ident()
                self.expect("=")
                fields[f] = self.parse_layout_init(consts)
                self.accept(",")
            return {"t": name, "f": fields}
        items = []
        while not self.accept("}"):
            items.append(self.parse_layout_init(consts))
            self.accept(",")
        return {"t": name, "a": items}

    def parse_layouts(self):
        consts = {}
        while self.peek()[0] != "eof":
            self.expect("constexpr")
            ty = self.parse_type()
            name = self.ident()
            self.expect("=")
            value = self.parse_layout_init(consts)
            self.expect(";")
            consts[name] = {"type": ty, "value": value}
        return consts


def flatten_layout(value, ty, types):
    """column numbers of a layout value in declaration order of its type (Reg = one column)"""
    if ty == "Reg":
        assert isinstance(value, int), (ty, value)
        return [value]
    kind = types[ty]
    if kind[0] == "struct":
        out = []
        for fname, fty in kind[1]:
            out += flatten_layout(value["f"][fname], fty, types)
        return out
    out = []
    assert len(value["a"]) == kind[2]
    for item in value["a"]:
        out += flatten_layout(item, kind[1], types)
    return out


def build(ref):
    base = os.path.join(ref, "risc0", "circuit", "rv32im-sys", "kernels", "cxx")
    types = Parser(tokenize(open(os.path.join(base, "types.h.inc")).read())).parse_types()
    consts = Parser(tokenize(open(os.path.join(base, "layout.cpp.inc")).read())).parse_layouts()
    text = open(os.path.join(base, "steps.cpp")).read()
    text = re.sub(r"^#include.*$", "", text, flags=re.M)
    funcs = Parser(tokenize(text)).parse_functions()
    layouts = {}
    for name in ("kLayout_Top", "kLayout_TopAccum", "kLayoutGlobal", "kLayoutMix"):
        c = consts[name]
        layouts[name] = {"type": c["type"], "cols": flatten_layout(c["value"], c["type"], types)}
    # keep only what step_Top / step_TopAccum reach
    reach, todo = set(), ["step_Top", "step_TopAccum"]

    def walk(node):
        if isinstance(node, list):
            if len(node) >= 2 and node[0] == "call" and isinstance(node[1], str) and node[1] in funcs:
                todo.append(node[1])
            for x in node:
                walk(x)
        elif isinstance(node, dict):
            for x in node.values():
                walk(x)

    while todo:
        f = todo.pop()
        if f in reach:
            continue
        reach.add(f)
        walk(funcs[f]["body"])
    funcs = {k: v for k, v in funcs.items() if k in reach}
    defs = dict(re.findall(r"constexpr size_t (\w+) = (\d+);", open(os.path.join(base, "defs.cpp.inc")).read()))
    return {"circuit": "rv32im", "source": "risc0/circuit/rv32im-sys/kernels/cxx/{steps.cpp,types.h.inc,layout.cpp.inc}",
            "regcounts": {k: int(v) for k, v in defs.items()}, "types": types, "layouts": layouts, "funcs": funcs}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    ap.add_argument("-o", default=os.path.join(root, "risc0_b200", "circuits", "rv32im_witgen.ir.json.gz"))
    a = ap.parse_args()
    sys.setrecursionlimit(100000)
    ir = build(a.ref)
    with gzip.GzipFile(a.o, "wb", mtime=0) as f:
        f.write(json.dumps(ir, separators=(",", ":")).encode())
    print("functions %d, types %d, layout columns %s -> %s (%d bytes)" % (
        len(ir["funcs"]), len(ir["types"]), {k: len(v["cols"]) for k, v in ir["layouts"].items()}, a.o, os.path.getsize(a.o)))


if __name__ == "__main__":
    main()
