"""A small SPIR-V interpreter: executes the reference's OWN compiled shaders on the CPU.

TEST INFRASTRUCTURE (like everything under oracle/): only tests/ uses it (the pin tests, the fixture generator
tests/golden/make_spirv_vectors.py and the verification scripts under tests/tools/).  The reference ships the binaries its engine loads
(ref: Assets/Compiled/Tracer.comp.spv, loaded at Source/GraphicsDevice.cpp:1091; Raytracer.comp.spv;
Fullscreen.frag.spv, Fullscreen.vert.spv, :1086), produced by glslangValidator from Assets/*.comp (ref: Assets/Compile.sh).  No Vulkan
driver exists in this image, so this module is the one way to *run the reference itself* here: it walks the
unoptimised, structured SPIR-V glslang emits, one invocation at a time.

Arithmetic: every OpF* is one IEEE-754 binary32 operation (computed in binary64 and rounded once, which is
exact for + - * / sqrt); nothing is contracted; GLSL.std.450 Sin / Cos / Pow go through libm in binary64 and are
rounded to binary32 -- one legal execution of the shader, chosen independently of the oracle's own arithmetic
contract (which fuses a few multiply-adds and uses its own polynomials), so comparisons against the oracle carry a
stated tolerance, and identity where the reference's arithmetic leaves no freedom (hit / miss decisions away
from ties, integer results, control flow).

Supported: exactly the instruction subset that occurs in the four binaries (checked at load time).
Functions can be called individually by their OpName (e.g. "trace_ray(struct-Ray-vf3-vf31;struct-Intersect-...;")
and calls to a named function can be intercepted (`hooks`), which is how the fixture generator substitutes the
repository's integer RNG for the shader's float hash rand() (DESIGN.md "RNG").
"""
import math
import struct

_F = struct.Struct("<f")
_I = struct.Struct("<I")


def f32(x):
    """Round a Python float to binary32 (overflow -> inf)."""
    try:
        return _F.unpack(_F.pack(x))[0]
    except OverflowError:
        return math.copysign(math.inf, x)


def _fdiv(a, b):
    if b == 0.0:
        if a != a or a == 0.0:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)
    if math.isinf(a) and math.isinf(b):
        return math.nan
    return f32(a / b)


def _fmul(a, b):
    if (math.isinf(a) and b == 0.0) or (math.isinf(b) and a == 0.0):
        return math.nan
    return f32(a * b)


def _fadd(a, b):
    return f32(a + b)          # inf + -inf -> nan by Python's float semantics


def _fsub(a, b):
    return f32(a - b)


def _sqrt(a):
    if a != a or a < 0.0:
        return math.nan
    if math.isinf(a):
        return a
    return f32(math.sqrt(a))


def _sin(a):
    return math.nan if (a != a or math.isinf(a)) else f32(math.sin(a))


def _cos(a):
    return math.nan if (a != a or math.isinf(a)) else f32(math.cos(a))


def _pow(a, b):
    # GLSL pow(x, y): undefined for x < 0, and for x == 0 with y <= 0; drivers evaluate exp2(y * log2(x))
    if a != a or b != b or a < 0.0:
        return math.nan
    if a == 0.0:
        return 0.0 if b > 0.0 else (math.nan if b == 0.0 else math.inf)
    try:
        return f32(math.pow(a, b))
    except OverflowError:
        return math.inf


class Pointer:
    """A SPIR-V pointer: a variable's cell plus an index path into its composite value."""
    __slots__ = ("cell", "path")

    def __init__(self, cell, path=()):
        self.cell, self.path = cell, path

    def load(self):
        v = self.cell[0]
        for i in self.path:
            v = v[i]
        return _copy(v)

    def store(self, value):
        value = _copy(value)
        if not self.path:
            self.cell[0] = value
            return
        v = self.cell[0]
        for i in self.path[:-1]:
            v = v[i]
        v[self.path[-1]] = value


def _copy(v):
    return [_copy(x) for x in v] if isinstance(v, list) else v


def _map1(fn, a):
    return [fn(x) for x in a] if isinstance(a, list) else fn(a)


def _map2(fn, a, b):
    if isinstance(a, list):
        return [fn(x, y) for x, y in zip(a, b)]
    return fn(a, b)


def _u32(x):
    return x & 0xFFFFFFFF


def _s32(x):
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


class Image:
    """A storage image: imageStore writes land in `texels[(x, y)]` as 4 floats (what the rgba8 unorm conversion sees)."""

    def __init__(self, width, height):
        self.size = [width, height]
        self.texels = {}


class Sampler2D:
    """A combined image sampler for Fullscreen.frag.spv: `fn(u, v)` returns the 4-float texel."""

    def __init__(self, fn):
        self.fn = fn


class Module:
    OPS_SEEN = set()

    def __init__(self, path):
        data = open(path, "rb").read()
        words = struct.unpack("<%dI" % (len(data) // 4), data)
        assert words[0] == 0x07230203, "not a SPIR-V binary"
        self.bound = words[3]
        self.names, self.member_names = {}, {}
        self.types, self.consts = {}, {}
        self.decor, self.member_decor = {}, {}
        self.globals = {}            # id -> (storage class, pointee type id)
        self.functions = {}          # id -> dict(params, blocks, order, first)
        self.entry = None
        self.glsl_ext = None
        cur = None
        i = 5
        while i < len(words):
            wc, op = words[i] >> 16, words[i] & 0xFFFF
            w = words[i + 1:i + wc]
            i += wc
            Module.OPS_SEEN.add(op)
            if op == 5:
                self.names[w[0]] = self._str(w[1:])
            elif op == 6:
                self.member_names[(w[0], w[1])] = self._str(w[2:])
            elif op == 11:
                assert self._str(w[1:]) == "GLSL.std.450"
                self.glsl_ext = w[0]
            elif op == 15:
                self.entry = w[1]
            elif op == 71:
                self.decor.setdefault(w[0], {})[w[1]] = w[2:]
            elif op == 72:
                self.member_decor.setdefault((w[0], w[1]), {})[w[2]] = w[3:]
            elif op == 19:
                self.types[w[0]] = ("void",)
            elif op == 20:
                self.types[w[0]] = ("bool",)
            elif op == 21:
                self.types[w[0]] = ("int", w[1], w[2])
            elif op == 22:
                self.types[w[0]] = ("float", w[1])
            elif op == 23:
                self.types[w[0]] = ("vector", w[1], w[2])
            elif op == 25:
                self.types[w[0]] = ("image",)
            elif op == 27:
                self.types[w[0]] = ("sampled_image",)
            elif op == 28:
                self.types[w[0]] = ("array", w[1], w[2])        # length is a constant id
            elif op == 29:
                self.types[w[0]] = ("runtime_array", w[1])
            elif op == 30:
                self.types[w[0]] = ("struct", list(w[1:]))
            elif op == 32:
                self.types[w[0]] = ("pointer", w[1], w[2])
            elif op == 33:
                self.types[w[0]] = ("function", w[1], list(w[2:]))
            elif op == 41:
                self.consts[w[1]] = True
            elif op == 42:
                self.consts[w[1]] = False
            elif op == 43:
                t = self.types[w[0]]
                if t[0] == "float":
                    self.consts[w[1]] = _F.unpack(_I.pack(w[2]))[0]
                else:
                    self.consts[w[1]] = _s32(w[2]) if t[2] else w[2]
            elif op == 44:
                self.consts[w[1]] = [self.consts[c] for c in w[2:]]
            elif op == 54:
                cur = {"id": w[1], "type": w[3], "params": [], "code": [], "labels": {}}
                self.functions[w[1]] = cur
            elif op == 55:
                cur["params"].append(w[1])
            elif op == 56:
                cur = None
            elif op == 59 and cur is None:
                self.globals[w[1]] = (w[2], self.types[w[0]][2], w[3] if len(w) > 3 else None)
            elif cur is not None:
                if op == 248:
                    cur["labels"][w[0]] = len(cur["code"])
                cur["code"].append((op, w))
            # everything else at module scope (OpSource, OpCapability, ...) carries no semantics here
        self.by_name = {n: fid for fid, n in self.names.items() if fid in self.functions}

    @staticmethod
    def _str(ws):
        b = b"".join(_I.pack(x) for x in ws)
        return b.split(b"\0", 1)[0].decode()

    def function(self, prefix):
        """The id of the function whose OpName starts with `prefix` (glslang mangles the signature after '(')."""
        hits = [fid for n, fid in self.by_name.items() if n.startswith(prefix)]
        assert len(hits) == 1, "%d functions match %r" % (len(hits), prefix)
        return hits[0]

    def zero(self, tid):
        t = self.types[tid]
        k = t[0]
        if k == "float":
            return 0.0
        if k == "int":
            return 0
        if k == "bool":
            return False
        if k == "vector":
            return [self.zero(t[1]) for _ in range(t[2])]
        if k == "array":
            return [self.zero(t[1]) for _ in range(self.consts[t[2]])]
        if k == "struct":
            return [self.zero(m) for m in t[1]]
        if k == "runtime_array":
            return []
        if k in ("image", "sampled_image"):
            return None
        raise NotImplementedError(k)


class Machine:
    """One shader invocation's state over a Module: global variables (push constants, buffers, images, built-ins)
    persist across calls; `run(fid, args)` executes one function."""

    def __init__(self, module, hooks=None):
        self.m = module
        self.g = {}                                   # global id -> cell
        for gid, (_, tid, init) in module.globals.items():
            self.g[gid] = [module.zero(tid) if init is None else _copy(module.consts[init])]
        self.hooks = {}                               # function id -> python callable(machine, args, call-site id) -> value
        for prefix, fn in (hooks or {}).items():
            self.hooks[module.function(prefix)] = fn
        self.steps = 0

    def global_cell(self, name):
        for gid in self.g:
            if self.m.names.get(gid) == name:
                return self.g[gid]
        raise KeyError(name)

    def global_by_builtin(self, builtin):
        for gid in self.g:
            d = self.m.decor.get(gid, {})
            if 11 in d and d[11][0] == builtin:
                return self.g[gid]
        raise KeyError(builtin)

    def global_by_binding(self, binding):
        for gid in self.g:
            d = self.m.decor.get(gid, {})
            if 33 in d and d[33][0] == binding:
                return self.g[gid]
        raise KeyError(binding)

    def global_by_storage(self, storage):
        for gid, (sc, _, _) in self.m.globals.items():
            if sc == storage:
                return self.g[gid]
        raise KeyError(storage)

    # ------------------------------------------------------------------------------------------
    def run(self, fid, args=()):
        m = self.m
        fn = m.functions[fid]
        code, labels = fn["code"], fn["labels"]
        v = {}                                        # result id -> value
        for pid, a in zip(fn["params"], args):
            v[pid] = a
        consts, g = m.consts, self.g

        def val(i):
            if i in v:
                return v[i]
            if i in consts:
                return consts[i]
            if i in g:
                return Pointer(g[i])
            raise KeyError("id %d (%s)" % (i, m.names.get(i)))

        pc, cur_label, prev_label = 0, None, None
        while True:
            op, w = code[pc]
            pc += 1
            self.steps += 1
            if op == 248:                                             # OpLabel
                prev_label, cur_label = cur_label, w[0]
            elif op == 61:                                            # OpLoad
                v[w[1]] = val(w[2]).load()
            elif op == 62:                                            # OpStore
                val(w[0]).store(val(w[1]))
            elif op == 65:                                            # OpAccessChain
                base = val(w[2])
                v[w[1]] = Pointer(base.cell, base.path + tuple(val(i) for i in w[3:]))
            elif op == 59:                                            # OpVariable (Function storage)
                cell = [m.zero(m.types[w[0]][2])]
                if len(w) > 3:
                    cell[0] = _copy(val(w[3]))
                v[w[1]] = Pointer(cell)
            elif op == 129:
                v[w[1]] = _map2(_fadd, val(w[2]), val(w[3]))
            elif op == 131:
                v[w[1]] = _map2(_fsub, val(w[2]), val(w[3]))
            elif op == 133:
                v[w[1]] = _map2(_fmul, val(w[2]), val(w[3]))
            elif op == 136:
                v[w[1]] = _map2(_fdiv, val(w[2]), val(w[3]))
            elif op == 127:
                v[w[1]] = _map1(lambda x: -x, val(w[2]))
            elif op == 142:                                           # OpVectorTimesScalar
                s = val(w[3])
                v[w[1]] = [_fmul(x, s) for x in val(w[2])]
            elif op == 148:                                           # OpDot: products summed left to right
                a, b = val(w[2]), val(w[3])
                acc = _fmul(a[0], b[0])
                for x, y in zip(a[1:], b[1:]):
                    acc = _fadd(acc, _fmul(x, y))
                v[w[1]] = acc
            elif op == 128:                                           # OpIAdd
                t = m.types[w[0]]
                r = _map2(lambda x, y: x + y, val(w[2]), val(w[3]))
                v[w[1]] = _map1(_s32 if (t[0] == "int" and t[2]) else _u32, r) if t[0] != "vector" else \
                    _map1(_s32 if m.types[t[1]][2] else _u32, r)
            elif op == 80:                                            # OpCompositeConstruct (vectors may be concatenated)
                t = m.types[w[0]]
                parts = [val(i) for i in w[2:]]
                if t[0] == "vector":
                    out = []
                    for p in parts:
                        out.extend(p) if isinstance(p, list) else out.append(p)
                    v[w[1]] = out
                else:
                    v[w[1]] = [_copy(p) for p in parts]
            elif op == 81:                                            # OpCompositeExtract (literal indices)
                x = val(w[2])
                for i in w[3:]:
                    x = x[i]
                v[w[1]] = _copy(x)
            elif op == 79:                                            # OpVectorShuffle
                a, b = val(w[2]), val(w[3])
                ab = a + b
                v[w[1]] = [ab[i] if i != 0xFFFFFFFF else 0.0 for i in w[4:]]
            elif op == 12:                                            # OpExtInst GLSL.std.450
                v[w[1]] = self._ext(w[3], [val(i) for i in w[4:]])
            elif op == 57:                                            # OpFunctionCall
                callee = w[2]
                a = [val(i) for i in w[3:]]
                v[w[1]] = self.hooks[callee](self, a, w[1]) if callee in self.hooks else self.run(callee, a)
            elif op == 254:                                           # OpReturnValue
                return val(w[0])
            elif op == 253:                                           # OpReturn
                return None
            elif op == 249:                                           # OpBranch
                pc = labels[w[0]]
            elif op == 250:                                           # OpBranchConditional
                pc = labels[w[1]] if val(w[0]) else labels[w[2]]
            elif op == 251:                                           # OpSwitch
                sel = val(w[0])
                target = w[1]
                for k in range(2, len(w), 2):
                    if _u32(sel) == w[k]:
                        target = w[k + 1]
                        break
                pc = labels[target]
            elif op in (246, 247):                                    # OpLoopMerge / OpSelectionMerge: structure only
                pass
            elif op == 245:                                           # OpPhi
                for k in range(2, len(w), 2):
                    if w[k + 1] == prev_label:
                        v[w[1]] = val(w[k])
                        break
                else:
                    raise RuntimeError("OpPhi: no incoming edge from %r" % prev_label)
            elif op == 184:
                v[w[1]] = _map2(lambda x, y: x < y, val(w[2]), val(w[3]))
            elif op == 186:
                v[w[1]] = _map2(lambda x, y: x > y, val(w[2]), val(w[3]))
            elif op == 188:
                v[w[1]] = _map2(lambda x, y: x <= y, val(w[2]), val(w[3]))
            elif op == 190:
                v[w[1]] = _map2(lambda x, y: x >= y, val(w[2]), val(w[3]))
            elif op == 180:
                v[w[1]] = _map2(lambda x, y: x == y, val(w[2]), val(w[3]))
            elif op == 182:
                v[w[1]] = _map2(lambda x, y: (x == x and y == y and x != y), val(w[2]), val(w[3]))
            elif op == 183:                                           # FUnordNotEqual
                v[w[1]] = _map2(lambda x, y: x != y, val(w[2]), val(w[3]))
            elif op in (170, 164):                                    # IEqual / LogicalEqual
                v[w[1]] = _map2(lambda x, y: x == y, val(w[2]), val(w[3]))
            elif op in (171, 165):
                v[w[1]] = _map2(lambda x, y: x != y, val(w[2]), val(w[3]))
            elif op == 176:                                           # ULessThan
                v[w[1]] = _map2(lambda x, y: _u32(x) < _u32(y), val(w[2]), val(w[3]))
            elif op == 177:                                           # SLessThan
                v[w[1]] = _map2(lambda x, y: _s32(x) < _s32(y), val(w[2]), val(w[3]))
            elif op == 178:                                           # ULessThanEqual
                v[w[1]] = _map2(lambda x, y: _u32(x) <= _u32(y), val(w[2]), val(w[3]))
            elif op == 179:
                v[w[1]] = _map2(lambda x, y: _s32(x) <= _s32(y), val(w[2]), val(w[3]))
            elif op == 172:
                v[w[1]] = _map2(lambda x, y: _u32(x) > _u32(y), val(w[2]), val(w[3]))
            elif op == 173:
                v[w[1]] = _map2(lambda x, y: _s32(x) > _s32(y), val(w[2]), val(w[3]))
            elif op == 166:
                v[w[1]] = _map2(lambda x, y: x or y, val(w[2]), val(w[3]))
            elif op == 167:
                v[w[1]] = _map2(lambda x, y: x and y, val(w[2]), val(w[3]))
            elif op == 168:
                v[w[1]] = _map1(lambda x: not x, val(w[2]))
            elif op == 154:
                v[w[1]] = any(val(w[2]))
            elif op == 155:
                v[w[1]] = all(val(w[2]))
            elif op == 169:                                           # OpSelect
                c, a, b = val(w[2]), val(w[3]), val(w[4])
                v[w[1]] = [x if k else y for k, x, y in zip(c, a, b)] if isinstance(c, list) else (a if c else b)
            elif op == 112:                                           # OpConvertUToF
                v[w[1]] = _map1(lambda x: f32(float(_u32(x))), val(w[2]))
            elif op == 111:                                           # OpConvertSToF
                v[w[1]] = _map1(lambda x: f32(float(_s32(x))), val(w[2]))
            elif op == 109:                                           # OpConvertFToU
                v[w[1]] = _map1(lambda x: _u32(int(x)) if x == x and not math.isinf(x) and x > 0 else 0, val(w[2]))
            elif op == 110:                                           # OpConvertFToS
                v[w[1]] = _map1(lambda x: _s32(int(x)) if x == x and not math.isinf(x) else 0, val(w[2]))
            elif op == 124:                                           # OpBitcast (same-width int <-> int here: uvec -> ivec)
                st, dt = None, m.types[w[0]]
                x = val(w[2])
                et = m.types[dt[1]] if dt[0] == "vector" else dt
                if et[0] == "int":
                    conv = _s32 if et[2] else _u32
                    v[w[1]] = _map1(lambda y: conv(y if isinstance(y, int) else _I.unpack(_F.pack(y))[0]), x)
                else:
                    v[w[1]] = _map1(lambda y: _F.unpack(_I.pack(_u32(y)))[0] if isinstance(y, int) else y, x)
            elif op in (194, 195, 196, 197, 198, 199):                # shifts and bitwise ops (Fullscreen.vert.spv)
                t = m.types[w[0]]
                et = m.types[t[1]] if t[0] == "vector" else t
                wrap = _s32 if et[2] else _u32
                fn = {194: lambda x, y: _u32(x) >> (y & 31), 195: lambda x, y: _s32(x) >> (y & 31), 196: lambda x, y: x << (y & 31),
                      197: lambda x, y: x | y, 198: lambda x, y: x ^ y, 199: lambda x, y: x & y}[op]
                v[w[1]] = _map1(wrap, _map2(fn, val(w[2]), val(w[3])))
            elif op == 68:                                            # OpArrayLength
                v[w[1]] = len(val(w[2]).load()[w[3]])
            elif op == 104:                                           # OpImageQuerySize
                v[w[1]] = list(val(w[2]).size)
            elif op == 99:                                            # OpImageWrite
                img, coord, texel = val(w[0]), val(w[1]), val(w[2])
                img.texels[(coord[0], coord[1])] = list(texel)
            elif op == 87:                                            # OpImageSampleImplicitLod
                smp, coord = val(w[2]), val(w[3])
                v[w[1]] = list(smp.fn(coord[0], coord[1]))
            else:
                raise NotImplementedError("SPIR-V opcode %d" % op)

    # ------------------------------------------------------------------------------------------
    def _ext(self, inst, a):
        if inst == 4:
            return _map1(abs, a[0])
        if inst == 6:
            return _map1(lambda x: (1.0 if x > 0.0 else (-1.0 if x < 0.0 else 0.0)), a[0])
        if inst == 10:                                                # Fract: x - floor(x)
            return _map1(lambda x: _fsub(x, float(math.floor(x))) if (x == x and not math.isinf(x)) else math.nan, a[0])
        if inst == 13:
            return _map1(_sin, a[0])
        if inst == 14:
            return _map1(_cos, a[0])
        if inst == 26:
            return _map2(_pow, a[0], a[1])
        if inst == 31:
            return _map1(_sqrt, a[0])
        if inst == 32:
            return _map1(lambda x: _fdiv(1.0, _sqrt(x)), a[0])
        if inst == 37:                                                # FMin: NaN operand -> the other one
            return _map2(lambda x, y: y if x != x else (x if y != y else min(x, y)), a[0], a[1])
        if inst == 40:                                                # FMax
            return _map2(lambda x, y: y if x != x else (x if y != y else max(x, y)), a[0], a[1])
        if inst == 43:                                                # FClamp = min(max(x, lo), hi)
            def clamp(x, lo, hi):
                t = lo if x != x else max(x, lo)
                return min(t, hi)
            if isinstance(a[0], list):
                return [clamp(x, lo, hi) for x, lo, hi in zip(a[0], a[1], a[2])]
            return clamp(a[0], a[1], a[2])
        if inst == 46:                                                # FMix = x * (1 - a) + y * a
            def mix(x, y, t):
                return _fadd(_fmul(x, _fsub(1.0, t)), _fmul(y, t))
            if isinstance(a[0], list):
                return [mix(x, y, t) for x, y, t in zip(a[0], a[1], a[2])]
            return mix(a[0], a[1], a[2])
        if inst == 66:                                                # Length
            return _sqrt(self._dot(a[0], a[0])) if isinstance(a[0], list) else abs(a[0])
        if inst == 68:                                                # Cross
            x, y = a
            return [_fsub(_fmul(x[1], y[2]), _fmul(y[1], x[2])), _fsub(_fmul(x[2], y[0]), _fmul(y[2], x[0])),
                    _fsub(_fmul(x[0], y[1]), _fmul(y[0], x[1]))]
        if inst == 69:                                                # Normalize = x / length(x)
            ln = _sqrt(self._dot(a[0], a[0]))
            return [_fdiv(x, ln) for x in a[0]]
        if inst == 71:                                                # Reflect = I - 2 * dot(N, I) * N
            i_, n = a
            d2 = _fmul(2.0, self._dot(n, i_))
            return [_fsub(x, _fmul(d2, y)) for x, y in zip(i_, n)]
        if inst == 72:                                                # Refract
            i_, n, eta = a
            d = self._dot(n, i_)
            k = _fsub(1.0, _fmul(_fmul(eta, eta), _fsub(1.0, _fmul(d, d))))
            if k < 0.0:
                return [0.0 for _ in i_]
            s = _fadd(_fmul(eta, d), _sqrt(k))
            return [_fsub(_fmul(eta, x), _fmul(s, y)) for x, y in zip(i_, n)]
        raise NotImplementedError("GLSL.std.450 instruction %d" % inst)

    @staticmethod
    def _dot(a, b):
        acc = _fmul(a[0], b[0])
        for x, y in zip(a[1:], b[1:]):
            acc = _fadd(acc, _fmul(x, y))
        return acc


def unorm8(x):
    """The rgba8 image store conversion: clamp to [0, 1], scale, round to nearest (NaN -> 0)."""
    if x != x:
        return 0
    x = min(max(x, 0.0), 1.0)
    return int(math.floor(f32(f32(x * 255.0) + 0.5)))


# ---- harness helpers for the two compute shaders ------------------------------------------------------
def set_frame_data(machine, aspect_ratio, seed, light_pos, cam_pos, cam_dir, cam_right, cam_up):
    """Fills the push-constant block (ref: Include/GraphicsDevice.h:20-29) member by member, by OpMemberName:
    Tracer.comp declares {aspect_ratio, seed, light_pos, camera}, Raytracer.comp has no `seed` member."""
    m = machine.m
    gid = [g for g, (sc, _, _) in m.globals.items() if sc == 9][0]
    tid = m.globals[gid][1]
    v3 = lambda a: [f32(float(a[0])), f32(float(a[1])), f32(float(a[2]))]
    fields = {"aspect_ratio": f32(float(aspect_ratio)), "seed": f32(float(seed)), "light_pos": v3(light_pos),
              "camera": [v3(cam_pos), v3(cam_dir), v3(cam_right), v3(cam_up)]}
    machine.g[gid][0] = [fields[m.member_names[(tid, k)]] for k in range(len(m.types[tid][1]))]


def run_compute(machine, width, height, pixels, triangles=None):
    """Runs the entry point for the given (x, y) invocations on a width x height storage image (binding 0); the
    triangle SSBO (binding 1, Tracer.comp only) is `triangles` = [[v0, v1, v2], ...].  Returns {(x, y): [r, g, b, a]}."""
    m = machine.m
    img = Image(width, height)
    machine.global_by_binding(0)[0] = img
    if triangles is not None:
        machine.global_by_binding(1)[0] = [[[[f32(float(c)) for c in v] for v in t] for t in triangles]]
    gid = machine.global_by_builtin(28)
    for (x, y) in pixels:
        gid[0] = [x, y, 0]
        machine.run(m.entry)
    return img.texels


class TracerRng:
    """Substitutes the repository's counter-based RNG for Tracer.comp's float hash rand() (ref: Tracer.comp:221-234;
    stated deviation, DESIGN.md "RNG") while the reference binary runs: every OpFunctionCall of rand() is one static
    call site, and the sites are, in code order, Tracer.comp:453 (r2), :454 (phi), :468 / :469 (light sample), :541
    (dielectric pick), :547 (Russian roulette) inside radiance() and :590 (dither) inside main().  The draw for a site is
    u01(pixel, sample, 32 * depth + slot) with slot 0 = :453 and :541, 1 = :454, 2 = :547, 3 + 2l / 4 + 2l = light l;
    depth = Russian-roulette draws so far in this radiance() call, sample = radiance() calls so far for this pixel."""

    def __init__(self, module, rand_u01):
        self.m, self.rand_u01 = module, rand_u01        # rand_u01(pixel, sample, dim) -> float
        self.rand_fid, self.rad_fid = module.function("rand("), module.function("radiance(")
        sites = []
        for fid in (self.rad_fid, module.entry):
            sites += [w[1] for op, w in module.functions[fid]["code"] if op == 57 and w[2] == self.rand_fid]
        assert len(sites) == 7, "Tracer.comp.spv: expected 7 rand() call sites, found %d" % len(sites)
        self.site = dict(zip(sites, ("r2", "phi", "cosa", "lphi", "pick", "rr", "dither")))
        self.pixel, self.sample, self.depth, self.light = 0, -1, 0, 0
        self.radiance_sum = None                           # the shader's `accum` before :585, formed like :580
        self.primary_ray = None                            # the Ray main() hands to radiance() (Tracer.comp:574)

    def hooks(self):
        return {"rand(": self._rand, "radiance(": self._radiance}

    def begin_pixel(self, pixel):
        self.pixel, self.sample, self.radiance_sum = pixel, -1, [0.0, 0.0, 0.0]

    def _radiance(self, machine, args, site):
        self.sample += 1
        self.depth = 0
        if self.sample == 0:
            self.primary_ray = args[0].load()
        r = machine.run(self.rad_fid, args)
        self.radiance_sum = [_fadd(a, b) for a, b in zip(self.radiance_sum, r)]
        return r

    def _rand(self, machine, args, site):
        kind = self.site[site]
        if kind == "dither":
            return f32(self.rand_u01(self.pixel, 0xFFFFFFFF, 0))
        if kind == "r2":
            self.light = 0
        slot = {"r2": 0, "pick": 0, "phi": 1, "rr": 2, "cosa": 3 + 2 * self.light, "lphi": 4 + 2 * self.light}[kind]
        u = f32(self.rand_u01(self.pixel, self.sample, 32 * self.depth + slot))
        if kind == "lphi":
            self.light += 1
        if kind == "rr":
            self.depth += 1
        return u


def bilinear_sampler(img_u8):
    """texture() of an R8G8B8A8_UNORM image through the reference's sampler (ref: Source/GraphicsDevice.cpp:770-794:
    LINEAR filter, CLAMP_TO_BORDER, opaque black border, one mip level).  Texel centres at +0.5, binary32 weights."""
    h, w = len(img_u8), len(img_u8[0])

    def texel(i, j):
        if i < 0 or j < 0 or i >= w or j >= h:
            return [0.0, 0.0, 0.0, 1.0]
        p = img_u8[j][i]
        return [_fdiv(float(p[0]), 255.0), _fdiv(float(p[1]), 255.0), _fdiv(float(p[2]), 255.0), _fdiv(float(p[3]), 255.0)]

    def mix(a, b, t):
        return [_fadd(_fmul(x, _fsub(1.0, t)), _fmul(y, t)) for x, y in zip(a, b)]

    def fn(u, v):
        s, t = _fsub(_fmul(u, float(w)), 0.5), _fsub(_fmul(v, float(h)), 0.5)
        fs0, ft0 = float(math.floor(s)), float(math.floor(t))
        fx, fy = _fsub(s, fs0), _fsub(t, ft0)
        i0, j0 = int(fs0), int(ft0)
        a = mix(texel(i0, j0), texel(i0 + 1, j0), fx)
        b = mix(texel(i0, j0 + 1), texel(i0 + 1, j0 + 1), fx)
        return mix(a, b, fy)
    return Sampler2D(fn)


def tracer_hit_id(found, intersection):
    """(kind << 28) | index of the primitive a Tracer.comp.spv trace_ray() result names in the shader-constant scene
    (ref: Tracer.comp:186-211), in the repository's AOV convention: 0 miss, kind 1 triangle, 2 sphere, 3 plane."""
    if not found:
        return 0
    (albedo, emissive, roughness, metalness, mtype), _t, _p, n = intersection
    planes = {(0.0, 1.0, 0.0): 0, (0.0, -1.0, 0.0): 1, (1.0, 0.0, 0.0): 2, (0.0, 0.0, -1.0): 3, (-1.0, 0.0, 0.0): 4}
    if tuple(n) in planes and metalness in (0.7, f32(0.7), 0.0) and mtype == 0 and emissive[0] == 0.0 and roughness != 0.0:
        return (3 << 28) | planes[tuple(n)]
    if mtype == 1:
        return (2 << 28) | 0                      # glass
    if emissive[0] != 0.0:
        return (2 << 28) | 1                      # light
    if metalness == 1.0:                          # mirror: the triangle (N = the constant cross product) or sphere 2
        return (1 << 28) | 0 if tuple(n) == (0.0, 0.0, 200.0) else (2 << 28) | 2
    return (2 << 28) | 3                          # plastic


def set_tracer_scene(module, spheres=None, planes=None):
    """Tracer.comp keeps its scene in two Private arrays that main() fills from constants at its start
    (`Sphere spheres[4] = {...}`, `Plane planes[5] = {...}`, ref: Tracer.comp:196-211).  This replaces the VALUES of those
    two constants in the loaded module -- data only, no instruction is touched -- so the reference binary renders another
    4-sphere / 5-plane scene.  spheres: 4 x [material, [x, y, z], r]; planes: 5 x [material, [nx, ny, nz], len];
    material = [albedo[3], emissive[3], roughness, metalness, type]."""
    main = module.functions[module.entry]
    gid = {module.names.get(g): g for g in module.globals}
    for name, value, count in (("spheres", spheres, 4), ("planes", planes, 5)):
        if value is None:
            continue
        assert len(value) == count, "%s: the loop bound is compiled in (%d)" % (name, count)
        stores = [w[1] for op, w in main["code"] if op == 62 and w[0] == gid[name]]
        assert len(stores) == 1 and stores[0] in module.consts
        conv = lambda v: [conv(x) for x in v] if isinstance(v, (list, tuple)) else (f32(float(v)) if isinstance(v, float) else int(v))
        module.consts[stores[0]] = conv(value)
