#!/usr/bin/env python3
"""Emit this backend's rv32im witness-generation code from the committed circuit IR
(risc0_b200/circuits/rv32im_witgen.ir.json.gz, produced by tools/witgen_ir.py).

Output: risc0_b200/csrc/gen/witgen_rv32im.inc - value types, the flattened column-layout table and one function per
IR function, written against the small runtime in csrc/witgen_rt.cuh (Val / ExtVal / Arr, ld / st on the column-major
matrices, the preflight externs). csrc/witgen.cu includes it for the device build; tests/witgen_host_check.cpp includes
the same text in a host build to check the generator on the CPU against the reference's compiled witgen.

What the emission does differently from the zirgen-generated C++ it was derived from:
  * layouts are not C++ objects: a bound layout is one 32-bit word (buffer id << 24 | index into ONE flat uint16 column
    table); every LAYOUT_LOOKUP / LAYOUT_SUBSCRIPT is folded to a constant added to that word at generation time;
  * field constants are emitted in Montgomery form; source locations become integer site ids (strings kept in a
    host-side table for error messages only); log / assert externs are dropped;
  * functions below a size threshold are force-inlined, the rest are real calls (bounds nvcc's compile time and keeps the
    instruction footprint of the divergent mux arms small).
"""
import argparse
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..")
P = 15 * 2**27 + 1
BUF_IDS = {"data": 0, "accum": 1, "global": 2, "mix": 3}
OUT_PARAMS = os.environ.get("WITGEN_OUT_PARAMS", "1") == "1"   # non-inlined functions return aggregates through a reference
INLINE_LIMIT = int(os.environ.get("WITGEN_INLINE_LIMIT", "1500"))   # JSON size of a body below which it is force-inlined


def mont(n):
    return (n % P) * (1 << 32) % P


class Gen:
    def __init__(self, ir):
        self.ir = ir
        self.types = ir["types"]
        self.funcs = ir["funcs"]
        self.flat_cache = {}
        self.sites = []
        self.const_tables = {}     # tuple of Montgomery words -> table name (constant arrays hoisted out of the functions)
        self.layout_base = {}
        self.layout_cols = []
        for name, lay in ir["layouts"].items():
            self.layout_base[name] = len(self.layout_cols)
            self.layout_cols += lay["cols"]
        self.tmp = 0

    # ---- types
    def flat_size(self, ty):
        if ty == "Reg":
            return 1
        if ty not in self.flat_cache:
            kind = self.types[ty]
            if kind[0] == "struct":
                self.flat_cache[ty] = sum(self.flat_size(f[1]) for f in kind[1])
            else:
                self.flat_cache[ty] = kind[2] * self.flat_size(kind[1])
        return self.flat_cache[ty]

    def resolve(self, ty):
        """strip ::value_type"""
        if ty.endswith("::value_type"):
            base = self.types[ty[:-len("::value_type")]]
            assert base[0] == "array"
            return base[1]
        return ty

    def layout_of(self, ty):
        assert ty.startswith("BoundLayout<"), ty
        return self.resolve(ty[len("BoundLayout<"):-1])

    def ctype(self, ty):
        if ty.startswith("BoundLayout<"):
            return "BL"
        ty = self.resolve(ty)
        if ty in ("Val", "ExtVal", "void"):
            return ty
        if ty in ("Index",):
            return "uint32_t"
        if ty in ("MutableBuf", "GlobalBuf"):
            return "uint32_t"
        assert ty in self.types, ty
        return ty

    def is_aggregate(self, ty):
        ty = self.resolve(ty)
        return ty in self.types or ty == "ExtVal"

    def value_types_in_order(self):
        """struct / array types reachable from function signatures and declarations, dependencies first"""
        need, order = set(), []

        def visit(ty):
            if ty.startswith("BoundLayout<"):
                return
            ty = self.resolve(ty)
            if ty not in self.types or ty in need:
                return
            need.add(ty)
            kind = self.types[ty]
            if kind[0] == "struct":
                for _, fty in kind[1]:
                    visit(fty)
            else:
                visit(kind[1])
            order.append(ty)

        def walk(node):
            if isinstance(node, list):
                if node and node[0] == "decl":
                    visit(node[1])
                if node and node[0] in ("struct", "array"):
                    visit(node[1])
                for x in node:
                    walk(x)
            elif isinstance(node, dict):
                for p in node.get("params", []):
                    visit(p[0])
                for x in node.values():
                    walk(x)

        for f in self.funcs.values():
            visit(f["ret"])
            for p in f["params"]:
                visit(p[0])
            walk(f["body"])
        return order

    # ---- layout expressions: (c++ text, IR layout type)
    def lay(self, e, env):
        k = e[0]
        if k == "var":
            return e[1], env[e[1]]
        if k == "ll":
            base, ty = self.lay(e[1], env)
            off = 0
            for step in e[2]:
                kind = self.types[ty]
                if step[0] == "f":
                    assert kind[0] == "struct", (ty, step)
                    for fname, fty in kind[1]:
                        if fname == step[1]:
                            ty = fty
                            break
                        off += self.flat_size(fty)
                    else:
                        raise KeyError("%s has no field %s" % (ty, step[1]))
                else:
                    assert kind[0] == "array", (ty, step)
                    off += step[1] * self.flat_size(kind[1])
                    ty = kind[1]
            return ("(%s + %du)" % (base, off) if off else base), ty
        if k == "ls":
            base, ty = self.lay(e[1], env)
            kind = self.types[ty]
            assert kind[0] == "array" and e[2][0] == "num", e
            off = e[2][1] * self.flat_size(kind[1])
            return ("(%s + %du)" % (base, off) if off else base), kind[1]
        if k == "bind":
            name, buf = e[1], e[2]
            return "bind_layout(%du, %s)" % (self.layout_base[name], buf), self.ir["layouts"][name]["type"]
        raise ValueError("not a layout expression: %r" % (e,))

    def const_table(self, e):
        """name of the namespace-scope table for an array literal whose elements are all field constants, else None.
        Such literals used to be emitted as temporaries (`Val30Array{{...}}[i]`): an array built on the thread's stack
        and indexed dynamically; they now live in constant memory (30 stores + a local load less per use)."""
        if e[0] != "array" or not e[2] or any(x[0] != "val" for x in e[2]):
            return None
        key = tuple(mont(x[1]) for x in e[2])
        if key not in self.const_tables:
            self.const_tables[key] = "kConstArr%d" % len(self.const_tables)
        return self.const_tables[key]

    # ---- value expressions
    def expr(self, e, env):
        k = e[0]
        if k == "val":
            return "Val::raw(%du)" % mont(e[1])
        if k == "ext":
            return "ExtVal{{%s}}" % ", ".join("Val::raw(%du)" % mont(x) for x in e[1:5])
        if k == "num":
            return "%du" % e[1]
        if k == "var":
            return e[1]
        if k == "fld":
            return "%s.%s" % (self.expr(e[1], env), e[2])
        if k == "idx":
            t = self.const_table(e[1])
            if t is not None:     # constant array indexed in place: a load from the hoisted table, no local array
                return "Val::raw(%s[%s])" % (t, self.expr(e[2], env))
            return "%s[%s]" % (self.expr(e[1], env), self.expr(e[2], env))
        if k == "bin":
            return "(%s %s %s)" % (self.expr(e[2], env), e[1], self.expr(e[3], env))
        if k == "neg":
            return "(-%s)" % self.expr(e[1], env)
        if k == "load":
            return "ld(ctx, %s, %s)" % (self.lay(e[1], env)[0], self.expr(e[2], env))
        if k == "loadext":
            return "ldext(ctx, %s, %s)" % (self.lay(e[1], env)[0], self.expr(e[2], env))
        if k in ("ll", "ls", "bind"):
            return self.lay(e, env)[0]
        if k == "struct":
            if not e[2]:
                return "%s{}" % e[1]
            return "%s{%s}" % (e[1], ", ".join(self.expr(f[1], env) for f in e[2]))
        if k == "array":
            return "%s{{%s}}" % (e[1], ", ".join(self.expr(x, env) for x in e[2]))
        if k == "extern":
            return "ext_%s(ctx%s)" % (e[1], "".join(", " + self.expr(a, env) for a in e[2]))
        if k == "call":
            name = e[1]
            if name in self.funcs:
                params = self.funcs[name]["params"]
                args = []
                for (pty, _), a in zip(params, e[2]):
                    args.append(self.lay(a, env)[0] if pty.startswith("BoundLayout<") else self.expr(a, env))
                return "%s(ctx%s)" % (name, "".join(", " + a for a in args))
            if name == "inv_0" and len(e[2]) == 1 and e[2][0][0] == "idx" and self.const_table(e[2][0][1]) is not None:
                # the inverse of an element of a constant array (ToBits: inv(2^i) for every bit of every decomposed
                # value): a second constant table instead of a ~45-product field inversion per use
                src = e[2][0][1]
                key = tuple(mont(pow(x[1] % P, P - 2, P)) for x in src[2])
                if key not in self.const_tables:
                    self.const_tables[key] = "kConstArr%d" % len(self.const_tables)
                return "Val::raw(%s[%s])" % (self.const_tables[key], self.expr(e[2][0][2], env))
            return "%s(%s)" % (name, ", ".join(self.expr(a, env) for a in e[2]))
        raise ValueError("unhandled expression %r" % (e[:2],))

    # ---- statements
    def stmts(self, body, env, ind, out, ret_target=None):
        pad = "  " * ind
        for s in body:
            k = s[0]
            if k == "decl":
                ty, name, init = s[1], s[2], s[3]
                if ty.startswith("BoundLayout<"):
                    text, lty = self.lay(init, env)
                    env[name] = lty
                    out.append("%sconst BL %s = %s;" % (pad, name, text))
                elif init is not None and init[0] in ("map", "reduce"):
                    self.emit_map(ty, name, init, env, ind, out)
                elif init is None:
                    out.append("%s%s %s;" % (pad, self.ctype(ty), name))
                else:
                    out.append("%s%s %s = %s;" % (pad, self.ctype(ty), name, self.expr(init, env)))
            elif k == "unpack":
                self.tmp += 1
                t = "u%d" % self.tmp
                out.append("%sconst auto %s = %s;" % (pad, t, self.expr(s[2], env)))
                for i, nm in enumerate(s[1]):
                    out.append("%sconst Val %s = %s[%d];" % (pad, nm, t, i))
            elif k == "assign":
                out.append("%s%s = %s;" % (pad, s[1], self.expr(s[2], env)))
            elif k == "if":
                for i, (cond, blk) in enumerate(s[1]):
                    c = cond
                    if c[0] == "call" and c[1] == "to_size_t":
                        c = c[2][0]
                    out.append("%s%sif (nz(%s)) {" % (pad, "} else " if i else "", self.expr(c, env)))
                    self.stmts(blk, dict(env), ind + 1, out, ret_target)
                if s[2] is not None:
                    out.append("%s} else {" % pad)
                    self.stmts(s[2], dict(env), ind + 1, out, ret_target)
                out.append("%s}" % pad)
            elif k == "eqz":
                self.sites.append(s[2])
                out.append("%seqz(ctx, %s, %du);" % (pad, self.expr(s[1], env), len(self.sites) - 1))
            elif k == "store":
                out.append("%sst(ctx, %s, %s);" % (pad, self.lay(s[1], env)[0], self.expr(s[2], env)))
            elif k == "storeext":
                out.append("%sstext(ctx, %s, %s);" % (pad, self.lay(s[1], env)[0], self.expr(s[2], env)))
            elif k == "expr":
                e = s[1]
                if e[0] == "extern" and e[1] in ("assert", "log"):
                    continue
                out.append("%s%s;" % (pad, self.expr(e, env)))
            elif k == "return":
                if s[1] is None:
                    out.append("%sreturn;" % pad)
                elif ret_target is not None:   # out-parameter form of a non-inlined function (see generate)
                    out.append("%s{ %s = %s; return; }" % (pad, ret_target, self.expr(s[1], env)))
                else:
                    out.append("%sreturn %s;" % (pad, self.expr(s[1], env)))
            elif k == "unreachable":
                out.append("%sunreachable(ctx);" % pad)
            else:
                raise ValueError("unhandled statement %r" % (s[:2],))

    def emit_map(self, ty, name, init, env, ind, out):
        """T name = map(arr, layout, lambda(elem, layout_elem)) / reduce(arr, init, layout, lambda(acc, elem, layout_elem)):
        a loop over the array with the element layout advancing by its flattened size"""
        pad = "  " * ind
        is_map = init[0] == "map"
        arr = init[1]
        lay_e = init[2] if is_map else init[3]
        lam = init[-1]
        aty = self.resolve(ty) if is_map else None
        base, lty = self.lay(lay_e, env)
        kind = self.types[lty]
        assert kind[0] == "array", lty
        n, esz = kind[2], self.flat_size(kind[1])
        self.tmp += 1
        t = self.tmp
        params = lam["params"]
        lenv = dict(env)
        lenv[params[-1][1]] = kind[1]
        sig = ", ".join(("BL %s" % p[1]) if p[0].startswith("BoundLayout<") else
                        ("const %s& %s" % (self.ctype(p[0]), p[1])) for p in params)
        if is_map:
            ret_ty = self.ctype(self.types[aty][1])
        else:
            ret_ty = self.ctype(ty)
        out.append("%sauto fn%d = [&](%s) -> %s {" % (pad, t, sig, ret_ty))
        self.stmts(lam["body"], lenv, ind + 1, out)
        out.append("%s};" % pad)
        ct = self.const_table(arr)
        if ct is not None:
            out.append("%sconst ConstArr arr%d{%s};" % (pad, t, ct))
        else:
            out.append("%sconst auto arr%d = %s;" % (pad, t, self.expr(arr, env)))
        if is_map:
            assert self.types[aty][2] == n, (aty, n)
            out.append("%s%s %s;" % (pad, self.ctype(ty), name))
            out.append("%sfor (uint32_t i%d = 0; i%d < %du; i%d++) %s[i%d] = fn%d(arr%d[i%d], %s + i%d * %du);" % (
                pad, t, t, n, t, name, t, t, t, t, base, t, esz))
        else:
            out.append("%s%s %s = %s;" % (pad, self.ctype(ty), name, self.expr(init[2], env)))
            out.append("%sfor (uint32_t i%d = 0; i%d < %du; i%d++) %s = fn%d(%s, arr%d[i%d], %s + i%d * %du);" % (
                pad, t, t, n, t, name, t, name, t, t, base, t, esz))

    def signature(self, name, out_form=False):
        f = self.funcs[name]
        ps = []
        for ty, pn in f["params"]:
            c = self.ctype(ty)
            if c == "BL" or c in ("uint32_t", "Val"):
                ps.append("%s %s" % (c, pn))
            else:
                ps.append("const %s& %s" % (c, pn))
        if out_form:
            return "void %s_o(WCtx& ctx%s, %s& ret_o)" % (name, "".join(", " + p for p in ps), self.ctype(f["ret"]))
        return "%s %s(WCtx& ctx%s)" % (self.ctype(f["ret"]), name, "".join(", " + p for p in ps))

    def returns_by_out_param(self, name):
        """A function that is a real call (not force-inlined) and returns a struct / array hands its result back through
        a reference to the caller's variable instead of by value. By value such a result travels in the PTX call's
        return parameter (`.param .align 4 .b8 retval0[64]` for the 16 bit registers of ToBits_16), and that is where
        the exec kernel went wrong on the device when ptxas had to work under a tight register cap (8 / 10 blocks per
        SM) or at -O3: BitwiseAndU16 keeps the first ToBits_16 result live across the second call, both decompositions
        passed their own eqz checks, and the AND of the two came out with single bits of the FIRST result changed
        (tests/test_gpu_witgen.py [all_insn]; the host build of the same text is clean under ASan / UBSan and with
        pattern-initialised locals). With the result written through a pointer the calling convention carries only
        scalars and pointers."""
        f = self.funcs[name]
        size = len(json.dumps(f["body"]))
        return OUT_PARAMS and size > INLINE_LIMIT and self.ctype(f["ret"]) not in ("Val", "void", "uint32_t", "BL")

    def generate(self):
        out = ["// GENERATED by tools/gen_witgen.py from risc0_b200/circuits/rv32im_witgen.ir.json.gz - do not edit.",
               "// Included inside namespace r0wg after witgen_rt.cuh (device build: csrc/witgen.cu; host check build: tests/).",
               ""]
        out.append("#define R0_WG_LAYOUT_WORDS %du" % len(self.layout_cols))
        for name, base in self.layout_base.items():
            out.append("#define R0_WG_%s %du" % (name.upper(), base))
        for k, v in self.ir["regcounts"].items():
            out.append("#define R0_WG_%s %du" % (k.upper(), v))
        # first accum column of the 'machine' part (ffi.cpp:51: kUserAccumSplit = kLayout_TopAccum.columns[0].col)
        acc = self.ir["layouts"]["kLayout_TopAccum"]
        fields = self.types[acc["type"]][1]
        assert fields[1][0] == "columns"
        out.append("#define R0_WG_USER_ACCUM_SPLIT %du" % acc["cols"][self.flat_size(fields[0][1])])
        out.append("#define R0_WG_LAYOUT_DATA { \\")
        cols = self.layout_cols
        for i in range(0, len(cols), 32):
            out.append("  " + ", ".join(str(c) for c in cols[i:i + 32]) + ", \\")
        out.append("}")
        out.append("")
        out.append("#ifndef R0_WG_TABLES_ONLY")
        for ty in self.value_types_in_order():
            kind = self.types[ty]
            if kind[0] == "array":
                out.append("typedef Arr<%s, %d> %s;" % (self.ctype(kind[1]), kind[2], ty))
            else:
                out.append("struct %s {%s};" % (ty, "".join(" %s %s;" % (self.ctype(f[1]), f[0]) for f in kind[1]) + " "))
        out.append("")
        for name in self.funcs:
            out.append("WG_FN %s;" % self.signature(name))
            if self.returns_by_out_param(name):
                out.append("WG_FN %s;" % self.signature(name, out_form=True))
        out.append("")
        body = []
        for name, f in self.funcs.items():
            size = len(json.dumps(f["body"]))
            qual = "WG_INLINE" if size <= INLINE_LIMIT else "WG_NOINLINE"
            env = {pn: self.layout_of(ty) for ty, pn in f["params"] if ty.startswith("BoundLayout<")}
            if self.returns_by_out_param(name):
                body.append("WG_NOINLINE %s {" % self.signature(name, out_form=True))
                self.stmts(f["body"], env, 1, body, ret_target="ret_o")
                body.append("}")
                body.append("WG_INLINE %s {" % self.signature(name))   # call sites stay expressions
                body.append("  %s r;" % self.ctype(f["ret"]))
                body.append("  %s_o(ctx%s, r);" % (name, "".join(", " + pn for _, pn in f["params"])))
                body.append("  return r;")
                body.append("}")
                continue
            body.append("%s %s {" % (qual, self.signature(name)))
            self.stmts(f["body"], env, 1, body)
            body.append("}")
        # constant arrays of the step functions (Montgomery words), hoisted to constant memory
        out.append("#if defined(__CUDACC__)")
        out.append("#define WG_TABLE static __constant__ uint32_t")
        out.append("#else")
        out.append("#define WG_TABLE static const uint32_t")
        out.append("#endif")
        out.append("struct ConstArr {")
        out.append("  const uint32_t* p;")
        out.append("  WG_RT Val operator[](uint32_t i) const { return Val::raw(p[i]); }")
        out.append("};")
        for key, tname in self.const_tables.items():
            out.append("WG_TABLE %s[%d] = {%s};" % (tname, len(key), ", ".join("%du" % w for w in key)))
        out.append("")
        out.extend(body)
        out.append("#endif  // R0_WG_TABLES_ONLY")
        out.append("")
        out.append("#ifdef R0_WG_SITE_STRINGS")
        out.append("static const char* const kEqzSites[] = {")
        for s in self.sites:
            out.append('  "%s",' % s.replace("\\", "\\\\").replace('"', '\\"'))
        out.append("};")
        out.append("#endif")
        return "\n".join(out) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ir", default=os.path.join(ROOT, "risc0_b200", "circuits", "rv32im_witgen.ir.json.gz"))
    ap.add_argument("-o", default=os.path.join(ROOT, "risc0_b200", "csrc", "gen", "witgen_rv32im.inc"))
    a = ap.parse_args()
    sys.setrecursionlimit(100000)
    ir = json.load(gzip.open(a.ir))
    text = Gen(ir).generate()
    if os.path.exists(a.o) and open(a.o).read() == text:
        print("%s: unchanged" % a.o)   # keep the file and its time: the two step kernels need minutes to rebuild
        return
    with open(a.o, "w") as f:
        f.write(text)
    print("%s: %d lines" % (a.o, text.count("\n")))


if __name__ == "__main__":
    main()
