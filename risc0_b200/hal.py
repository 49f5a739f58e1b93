"""Host-side mirror of the reference `Hal` / `Buffer` traits (risc0/zkp/src/hal/mod.rs:39-258) over the r0b200 C ABI.

Method names, argument order and meaning follow the Rust trait so the parity tests read like the reference's own
`hal::testutil` A/B tests (hal/mod.rs:319-616). Host data are numpy uint32 arrays of Montgomery words
(Elem = 1 word, ExtElem = 4 words, Digest = 8 words).
"""
import ctypes as C

import numpy as np

from ._lib import R0B200Error, check, load_library

_u32p = C.POINTER(C.c_uint32)
POSEIDON2, SHA256 = 0, 1
P = 15 * 2**27 + 1
INVALID = 0xFFFFFFFF
CHECK_SIZE = 16


def _np_ptr(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u32p)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class _Alloc:
    """owns one device allocation; freed (stream-ordered) when the last Buffer view drops"""

    def __init__(self, hal, nbytes):
        self.hal = hal
        self.nbytes = nbytes
        p = C.c_void_p()
        check(hal._l.r0b200_alloc(hal._ctx, C.c_size_t(nbytes), C.byref(p)))
        self.ptr = p.value

    def __del__(self):
        try:
            if self.ptr and self.hal._ctx:
                self.hal._l.r0b200_free(self.hal._ctx, C.c_void_p(self.ptr))
        except Exception:
            pass


class Buffer:
    """trait Buffer<T> (hal/mod.rs:39-53): aliasing slices, get_at, view, to_vec. `words` = u32 words per element."""

    def __init__(self, hal, name, size, words, alloc=None, offset=0):
        self.hal, self._name, self._size, self.words, self.offset = hal, name, int(size), words, int(offset)
        self.alloc = alloc if alloc is not None else _Alloc(hal, self._size * words * 4)

    def name(self):
        return self._name

    def size(self):
        return self._size

    @property
    def ptr(self):
        return C.c_void_p(self.alloc.ptr + self.offset * self.words * 4)

    def slice(self, offset, size):
        assert offset + size <= self._size
        return Buffer(self.hal, self._name, size, self.words, self.alloc, self.offset + offset)

    def view(self):
        out = np.empty(self._size * self.words, dtype=np.uint32)
        check(self.hal._l.r0b200_copy_d2h(self.hal._ctx, _np_ptr(out), self.ptr, C.c_size_t(out.nbytes)))
        return out

    to_vec = view

    def get_at(self, idx):
        out = np.empty(self.words, dtype=np.uint32)
        src = C.c_void_p(self.alloc.ptr + (self.offset + idx) * self.words * 4)
        check(self.hal._l.r0b200_copy_d2h(self.hal._ctx, _np_ptr(out), src, C.c_size_t(out.nbytes)))
        return out

    def write(self, data):
        data = _u32(data)
        assert data.size == self._size * self.words, (data.size, self._size, self.words)
        check(self.hal._l.r0b200_copy_h2d(self.hal._ctx, self.ptr, _np_ptr(data), C.c_size_t(data.nbytes)))
        # the copy is stream-ordered from pageable memory: staged before the call returns


def _lg(n):
    lg = int(n).bit_length() - 1
    if n <= 0 or (1 << lg) != n:
        raise R0B200Error("size %d is not a power of two" % n)
    return lg


class B200Hal:
    """impl Hal for the B200 backend. One instance = one device ordinal + one stream (hal/cuda.rs:397-421 analogue)."""

    def __init__(self, device=0, hashfn="poseidon2"):
        self._l = load_library()
        self._ctx = C.c_void_p()
        self.hash = {"poseidon2": POSEIDON2, "sha-256": SHA256}[hashfn]
        self.hashfn = hashfn
        check(self._l.r0b200_create(int(device), C.byref(self._ctx)))

    def close(self):
        if self._ctx:
            self._l.r0b200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection
    def has_unified_memory(self):
        return False

    def sync(self):
        check(self._l.r0b200_sync(self._ctx))

    def launch_count(self):
        return int(self._l.r0b200_launch_count(self._ctx))

    def stream(self):
        return self._l.r0b200_stream(self._ctx)

    def timer_start(self):
        check(self._l.r0b200_timer_start(self._ctx))

    def timer_stop(self):
        ms = C.c_float()
        check(self._l.r0b200_timer_stop(self._ctx, C.byref(ms)))
        return ms.value

    def bytes_peak(self):
        return int(self._l.r0b200_bytes_peak(self._ctx))

    def profile_begin(self):
        check(self._l.r0b200_profile_begin(self._ctx))

    def profile_end(self):
        """{op: {"ms": device ms, "n": launches, "bytes": algorithmic bytes}} since profile_begin"""
        import json
        buf = C.create_string_buffer(1 << 16)
        check(self._l.r0b200_profile_end(self._ctx, buf, C.c_size_t(len(buf))))
        return json.loads(buf.value.decode())

    # ---- allocation (hal/mod.rs:67-100)
    def alloc_elem(self, name, size):
        return Buffer(self, name, size, 1)

    def alloc_extelem(self, name, size):
        return Buffer(self, name, size, 4)

    def alloc_digest(self, name, size):
        return Buffer(self, name, size, 8)

    def alloc_u32(self, name, size):
        return Buffer(self, name, size, 1)

    def alloc_elem_init(self, name, size, value_mont):
        b = self.alloc_elem(name, size)
        check(self._l.r0b200_fill_u32(self._ctx, b.ptr, C.c_uint32(int(value_mont)), C.c_size_t(size)))
        return b

    def alloc_extelem_zeroed(self, name, size):
        b = self.alloc_extelem(name, size)
        check(self._l.r0b200_fill_u32(self._ctx, b.ptr, C.c_uint32(0), C.c_size_t(size * 4)))
        return b

    def _copy_from(self, name, data, words):
        data = _u32(data)
        b = Buffer(self, name, data.size // words, words)
        b.write(data)
        return b

    def copy_from_elem(self, name, data):
        return self._copy_from(name, data, 1)

    def copy_from_extelem(self, name, data):
        return self._copy_from(name, data, 4)

    def copy_from_digest(self, name, data):
        return self._copy_from(name, data, 8)

    def copy_from_u32(self, name, data):
        return self._copy_from(name, data, 1)

    # ---- NTT family
    def batch_expand_into_evaluate_ntt(self, output, input, count, expand_bits):
        in_row, out_row = input.size() // count, output.size() // count
        if in_row * count != input.size() or out_row != in_row << expand_bits:
            raise R0B200Error("batch_expand_into_evaluate_ntt: size mismatch")
        check(self._l.r0b200_batch_expand_into_evaluate_ntt(self._ctx, output.ptr, input.ptr, C.c_size_t(count),
                                                            C.c_uint32(_lg(in_row)), C.c_uint32(expand_bits)))

    def batch_interpolate_ntt(self, io, count):
        row = io.size() // count
        assert row * count == io.size()
        check(self._l.r0b200_batch_interpolate_ntt(self._ctx, io.ptr, C.c_size_t(count), C.c_uint32(_lg(row))))

    def batch_interpolate_ntt_zk(self, io, count):
        """fused batch_interpolate_ntt + zk_shift (make_coeffs, prove/prover.rs:38-48)"""
        row = io.size() // count
        check(self._l.r0b200_batch_interpolate_ntt_zk(self._ctx, io.ptr, C.c_size_t(count), C.c_uint32(_lg(row))))

    def batch_bit_reverse(self, io, count):
        row = io.size() // count
        assert row * count == io.size()
        check(self._l.r0b200_batch_bit_reverse(self._ctx, io.ptr, C.c_size_t(count), C.c_uint32(_lg(row))))

    def zk_shift(self, io, count):
        row = io.size() // count
        check(self._l.r0b200_zk_shift(self._ctx, io.ptr, C.c_size_t(count), C.c_uint32(_lg(row))))

    def batch_evaluate_any(self, coeffs, poly_count, which, xs, out):
        n = coeffs.size() // poly_count
        assert n * poly_count == coeffs.size() and xs.size() == which.size() == out.size()
        check(self._l.r0b200_batch_evaluate_any(self._ctx, coeffs.ptr, C.c_size_t(poly_count), C.c_uint32(_lg(n)),
                                                which.ptr, xs.ptr, out.ptr, C.c_size_t(which.size())))

    def mix_poly_coeffs(self, out, mix_start, mix, input, combos, input_size, count):
        """`combos` is the host u32 array of combo ids (the reference uploads it with copy_from_u32 first)"""
        combos = _u32(combos)
        assert combos.size == input_size and input.size() == input_size * count
        check(self._l.r0b200_mix_poly_coeffs(self._ctx, out.ptr, _np_ptr(_u32(mix_start)), _np_ptr(_u32(mix)),
                                             input.ptr, _np_ptr(combos), C.c_size_t(input_size), C.c_size_t(count)))

    # ---- element-wise
    def eltwise_add_elem(self, output, input1, input2):
        assert output.size() == input1.size() == input2.size()
        check(self._l.r0b200_eltwise_add_elem(self._ctx, output.ptr, input1.ptr, input2.ptr, C.c_size_t(output.size())))

    def eltwise_sum_extelem(self, output, input):
        count = output.size() // 4
        to_add = input.size() // count
        assert output.size() == 4 * count and input.size() == count * to_add
        check(self._l.r0b200_eltwise_sum_extelem(self._ctx, output.ptr, input.ptr, C.c_size_t(count), C.c_size_t(to_add)))

    def eltwise_copy_elem(self, output, input):
        assert output.size() == input.size()
        check(self._l.r0b200_eltwise_copy_elem(self._ctx, output.ptr, input.ptr, C.c_size_t(output.size())))

    def eltwise_copy_elem_slice(self, into, from_host, from_rows, from_cols, from_offset, from_stride, into_offset,
                                into_stride):
        check(self._l.r0b200_eltwise_copy_elem_slice(self._ctx, into.ptr, _np_ptr(_u32(from_host)),
                                                     C.c_size_t(from_rows), C.c_size_t(from_cols),
                                                     C.c_size_t(from_offset), C.c_size_t(from_stride),
                                                     C.c_size_t(into_offset), C.c_size_t(into_stride)))

    def eltwise_zeroize_elem(self, elems):
        check(self._l.r0b200_eltwise_zeroize_elem(self._ctx, elems.ptr, C.c_size_t(elems.size())))

    def fri_fold(self, output, input, mix):
        count = output.size() // 4
        assert input.size() == output.size() * 16
        check(self._l.r0b200_fri_fold(self._ctx, output.ptr, input.ptr, C.c_size_t(count), _np_ptr(_u32(mix))))

    # ---- hashing
    def hash_rows(self, output, matrix):
        rows = output.size()
        cols = matrix.size() // rows if rows else 0
        assert matrix.size() == rows * cols
        check(self._l.r0b200_hash_rows(self._ctx, self.hash, output.ptr, matrix.ptr, C.c_size_t(rows), C.c_size_t(cols)))

    def hash_fold(self, io, input_size, output_size):
        assert io.size() >= 2 * input_size
        check(self._l.r0b200_hash_fold(self._ctx, self.hash, io.ptr, C.c_size_t(input_size), C.c_size_t(output_size)))

    def merkle_build(self, nodes, matrix, rows, cols):
        """MerkleTreeProver::new's hash_rows + all hash_fold levels in one call"""
        assert nodes.size() == 2 * rows and matrix.size() == rows * cols
        check(self._l.r0b200_merkle_build(self._ctx, self.hash, nodes.ptr, matrix.ptr, C.c_size_t(rows), C.c_size_t(cols)))

    # ---- gather / scatter / misc
    def gather_sample(self, dst, src, idx, size, stride):
        check(self._l.r0b200_gather_sample(self._ctx, dst.ptr, src.ptr, C.c_size_t(idx), C.c_size_t(size),
                                           C.c_size_t(stride)))

    def scatter(self, into, index, offsets, values):
        index, offsets, values = _u32(index), _u32(offsets), _u32(values)
        if index.size == 0:
            return
        check(self._l.r0b200_scatter(self._ctx, into.ptr, _np_ptr(index), C.c_size_t(index.size), _np_ptr(offsets),
                                     _np_ptr(values)))

    def prefix_products(self, io):
        check(self._l.r0b200_prefix_products(self._ctx, io.ptr, C.c_size_t(io.size())))

    def combos_prepare(self, combos, coeff_u, combo_count, cycles, reg_sizes, reg_combo_ids, mix):
        coeff_u, reg_sizes, reg_combo_ids = _u32(coeff_u), _u32(reg_sizes), _u32(reg_combo_ids)
        check(self._l.r0b200_combos_prepare(self._ctx, combos.ptr, _np_ptr(coeff_u), C.c_size_t(coeff_u.size // 4),
                                            C.c_uint32(combo_count), C.c_size_t(cycles), _np_ptr(reg_sizes),
                                            _np_ptr(reg_combo_ids), C.c_uint32(reg_sizes.size), _np_ptr(_u32(mix))))

    def combos_divide(self, combos, chunks, cycles):
        """chunks: list of (i, [pow ext elems]) as in hal/mod.rs:236-257; chunk order = position in `combos`"""
        pow_begin, pows = [0], []
        for _, ps in chunks:
            for p in ps:
                pows.extend(int(x) for x in p)
            pow_begin.append(len(pows) // 4)
        pows = _u32(pows if pows else [0, 0, 0, 0])
        check(self._l.r0b200_combos_divide(self._ctx, combos.ptr, C.c_size_t(len(chunks)), _np_ptr(_u32(pow_begin)),
                                           _np_ptr(pows), C.c_size_t(cycles)))

    # ---- CircuitHal (hal/mod.rs:265-290) for rv32im
    def eval_check_rv32im(self, check_buf, groups, globals_, poly_mix, po2, steps):
        """groups = [accum, code, data] evaluated buffers (tap-group order); globals_ = [mix, out] device buffers
        (the order rv32im/src/prove/hal/mod.rs:216 passes them)"""
        assert steps == 1 << po2 and check_buf.size() == 16 * steps
        accum, code, data = groups
        mix, out = globals_
        err = self._l.r0b200_eval_check_rv32im(self._ctx, check_buf.ptr, code.ptr, data.ptr, accum.ptr, mix.ptr,
                                                     out.ptr, _np_ptr(_u32(poly_mix)), C.c_uint32(po2))
        check(err)

    def eval_check_recursion(self, check_buf, groups, globals_, poly_mix, po2, steps):
        """recursion circuit: groups = [accum, ctrl, data] (tap-group order), globals_ = [mix, out]"""
        assert steps == 1 << po2 and check_buf.size() == 16 * steps
        accum, ctrl, data = groups
        mix, out = globals_
        check(self._l.r0b200_eval_check_recursion(self._ctx, check_buf.ptr, ctrl.ptr, data.ptr, accum.ptr, mix.ptr,
                                                  out.ptr, _np_ptr(_u32(poly_mix)), C.c_uint32(po2)))


class PreflightTraceStruct(C.Structure):
    """r0b200_preflight_trace = RawPreflightTrace (rv32im-sys/src/lib.rs:63-72) with host pointers"""
    _fields_ = [("cycles", C.c_void_p), ("txns", C.c_void_p), ("bigint_bytes", C.c_void_p), ("txns_len", C.c_uint32),
                ("bigint_bytes_len", C.c_uint32), ("table_split_cycle", C.c_uint32)]


def _trace_struct(pf):
    """pf: risc0_b200.preflight.PreflightResults (or anything with .cycles / .txns structured arrays, .bigint_bytes,
    .table_split_cycle). Returns (struct, keep-alive tuple)."""
    cycles = np.ascontiguousarray(pf.cycles)
    txns = np.ascontiguousarray(pf.txns)
    bigint = np.ascontiguousarray(pf.bigint_bytes, dtype=np.uint8)
    assert cycles.dtype.itemsize == 36 and txns.dtype.itemsize == 20
    st = PreflightTraceStruct(cycles.ctypes.data, txns.ctypes.data if len(txns) else None,
                              bigint.ctypes.data if len(bigint) else None, len(txns), len(bigint), int(pf.table_split_cycle))
    return st, (cycles, txns, bigint)


class DeviceTrace:
    """a preflight trace resident on the device (r0b200_trace_upload), sorted there by (major, minor)"""

    def __init__(self, hal, pf):
        self.hal = hal
        self.cycles = len(pf.cycles)
        st, keep = _trace_struct(pf)
        self._h = C.c_void_p()
        check(hal._l.r0b200_trace_upload(hal._ctx, C.byref(st), C.c_uint32(self.cycles), C.byref(self._h)))
        hal.sync()   # the host arrays may go away after this

    def close(self):
        # a trace belongs to its context: once the Hal is closed the handle is dangling and must not be touched
        if self._h and self.hal._ctx:
            self.hal._l.r0b200_trace_free(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Rv32imCircuitHal:
    """CircuitWitnessGenerator + CircuitAccumulator for rv32im (rv32im/src/prove/hal/mod.rs:82-102) on the device"""

    PARALLEL, SEQ_FORWARD, SEQ_REVERSE = 0, 1, 2

    def __init__(self, hal):
        self.hal = hal

    def generate_witness(self, mode, trace, global_buf, data_buf):
        check(self.hal._l.r0b200_witgen_rv32im(self.hal._ctx, C.c_uint32(mode), trace._h, global_buf.ptr, data_buf.ptr))

    def step_accum(self, trace, data_buf, accum_buf, global_buf, mix_buf):
        check(self.hal._l.r0b200_accum_rv32im(self.hal._ctx, trace._h, data_buf.ptr, accum_buf.ptr, global_buf.ptr,
                                              mix_buf.ptr))


class WitnessGenerator:
    """prove/witgen/mod.rs:90-224 over the device Hal: new() = INVALID fill + injector scatter + generate_witness +
    zeroize; accum(mix) = step_accum + zeroize. Buffers stay on the device."""

    N_CODE, N_DATA, N_ACCUM, N_GLOBAL, N_MIX = 1, 211, 103, 90, 36

    def __init__(self, hal, pf, mode=Rv32imCircuitHal.PARALLEL):
        self.hal, self.pf = hal, pf
        self.circuit = Rv32imCircuitHal(hal)
        self.cycles = len(pf.cycles)
        self.trace = DeviceTrace(hal, pf)
        self.global_ = hal.copy_from_elem("global", pf.global_)
        self.code = hal.alloc_elem_init("code", self.cycles * self.N_CODE, INVALID)
        self.data = hal.alloc_elem_init("data", self.cycles * self.N_DATA, INVALID)
        index, offsets, values = pf.injector
        hal.scatter(self.data, index, offsets, values)
        self.circuit.generate_witness(mode, self.trace, self.global_, self.data)
        hal.eltwise_zeroize_elem(self.global_)
        hal.eltwise_zeroize_elem(self.code)
        hal.eltwise_zeroize_elem(self.data)
        self.accum_buf = hal.alloc_elem_init("accum", self.cycles * self.N_ACCUM, INVALID)

    def accum(self, mix):
        # segments with bigint cycles: the BigIntAccumState cells depend on the mix and are scattered now
        # (witgen/mod.rs:186-207); every other segment has none
        inj = self.pf.bigint_accum_injector(mix) if getattr(self.pf, "has_bigint", False) else None
        if inj is not None:
            self.hal.scatter(self.accum_buf, *inj)
        mix_buf = self.hal.copy_from_elem("mix", mix)
        self.circuit.step_accum(self.trace, self.data, self.accum_buf, self.global_, mix_buf)
        self.hal.eltwise_zeroize_elem(self.accum_buf)
        return mix_buf


class SegmentProver:
    """Host-side mirror of `SegmentProver::prove_core`'s prove_inner block for rv32im
    (risc0/circuit/rv32im/src/prove/hal/mod.rs:171-222) over r0b200_prove_rv32im: commit code/data/accum, eval_check,
    DEEP, FRI -> seal. `witness` arrays are column-major Montgomery words (numpy = host memory, Buffer = device)."""

    def __init__(self, hal):
        self.hal = hal
        self.seal_cap = 1 << 20
        self._seal = np.zeros(self.seal_cap, dtype=np.uint32)
        self._roots = np.zeros(8 * 16, dtype=np.uint32)
        self._qpos = np.zeros(50, dtype=np.uint32)

    # (code/ctrl, data, accum) column counts and global words per circuit
    SHAPES = {"rv32im": (1, 211, 103, 90), "recursion": (23, 128, 12, 32)}

    def upload(self, po2, code, data, accum, circuit="rv32im"):
        """enqueue the host->device copy of a witness (numpy arrays, ideally views of pinned memory that stay alive until
        the proof has run) and return a handle for prove(uploaded=...). Call it for segment s+1 before proving segment s
        to hide the transfer behind compute (the reference's depth-2 queues, r0vm/src/actors/worker.rs:70-76)."""
        hal = self.hal
        n = 1 << po2
        c_code, c_data, c_accum, _ = self.SHAPES[circuit]
        code, data = _u32(code), _u32(data)
        assert code.size == c_code * n and data.size == c_data * n
        if accum is not None:
            accum = _u32(accum)
            assert accum.size == c_accum * n
        h = C.c_void_p()
        check(hal._l.r0b200_witness_upload(hal._ctx, C.c_int(0 if circuit == "rv32im" else 1), C.c_uint32(po2),
                                           _np_ptr(code), _np_ptr(data), _np_ptr(accum) if accum is not None else None,
                                           C.byref(h)))
        return (h, (code, data, accum))   # keep the host arrays alive with the handle

    def begin(self, po2, code, data, glob, circuit="rv32im", uploaded=None):
        """prove_core up to the mix draw (rv32im/src/prove/hal/mod.rs:181-213): commits code and data and returns
        (proof handle, mix words). The caller computes accum from the mix and calls finish()."""
        hal = self.hal
        n = 1 << po2
        c_code, c_data, _, n_glob = self.SHAPES[circuit]
        glob = _u32(glob)
        assert glob.size == n_glob
        mix = np.zeros(64, dtype=np.uint32)
        h = C.c_void_p()
        cid = C.c_int(0 if circuit == "rv32im" else 1)
        if uploaded is not None:
            uh, _keep = uploaded
            try:
                check(hal._l.r0b200_prove_begin(hal._ctx, cid, hal.hash, C.c_uint32(po2), None, None, C.c_int(1), uh,
                                                _np_ptr(glob), _np_ptr(mix), C.c_size_t(mix.size), C.byref(h)))
            finally:
                hal._l.r0b200_witness_free(uh)
        else:
            on_host = isinstance(data, np.ndarray)
            if on_host:
                code, data = _u32(code), _u32(data)
                assert code.size == c_code * n and data.size == c_data * n
                ptrs = [_np_ptr(code), _np_ptr(data)]
            else:
                assert code.size() == c_code * n and data.size() == c_data * n
                ptrs = [code.ptr, data.ptr]
            check(hal._l.r0b200_prove_begin(hal._ctx, cid, hal.hash, C.c_uint32(po2), ptrs[0], ptrs[1],
                                            C.c_int(1 if on_host else 0), None, _np_ptr(glob), _np_ptr(mix),
                                            C.c_size_t(mix.size), C.byref(h)))
        nmix = 36 if circuit == "rv32im" else 20
        return h, mix[:nmix].copy()

    def finish(self, proof, accum):
        """commit accum, eval_check, DEEP, FRI -> (seal, roots, query positions); consumes the handle"""
        hal = self.hal
        on_host = isinstance(accum, np.ndarray)
        if on_host:
            accum = _u32(accum)
        seal_len, nroots = C.c_size_t(0), C.c_size_t(0)
        check(hal._l.r0b200_prove_finish(proof, _np_ptr(accum) if on_host else accum.ptr, C.c_int(1 if on_host else 0),
                                         _np_ptr(self._seal), C.c_size_t(self.seal_cap), C.byref(seal_len),
                                         _np_ptr(self._roots), C.c_size_t(16), C.byref(nroots), _np_ptr(self._qpos)))
        return (self._seal[:seal_len.value].copy(), self._roots[:8 * nroots.value].reshape(-1, 8).copy(),
                self._qpos.copy())

    def abort(self, proof):
        self.hal._l.r0b200_prove_abort(proof)

    def upload_segment(self, pf):
        """r0b200_segment_upload: enqueue the copy of a PreflightResults (trace, injector, globals) on the copy stream.
        Returns a handle for prove_segment(); call it for segment s + 1 before proving segment s."""
        hal = self.hal
        st, keep = _trace_struct(pf)
        index, offsets, values = (_u32(a) for a in pf.injector)
        glob = _u32(pf.global_)
        h = C.c_void_p()
        check(hal._l.r0b200_segment_upload(hal._ctx, C.c_uint32(pf.po2), C.byref(st), _np_ptr(glob), _np_ptr(index),
                                           C.c_size_t(index.size), _np_ptr(offsets), _np_ptr(values), C.byref(h)))
        return (h, (keep, index, offsets, values, glob))

    def prove_segment(self, segment, free=True):
        """prove_core on an uploaded segment -> (seal, roots, query positions, globals)"""
        hal = self.hal
        h, _keep = segment
        glob_out = np.zeros(90, dtype=np.uint32)
        seal_len, nroots = C.c_size_t(0), C.c_size_t(0)
        try:
            check(hal._l.r0b200_prove_segment(hal._ctx, hal.hash, h, _np_ptr(self._seal), C.c_size_t(self.seal_cap),
                                              C.byref(seal_len), _np_ptr(self._roots), C.c_size_t(16), C.byref(nroots),
                                              _np_ptr(self._qpos), _np_ptr(glob_out)))
        finally:
            if free:
                hal._l.r0b200_segment_free(h)
        return (self._seal[:seal_len.value].copy(), self._roots[:8 * nroots.value].reshape(-1, 8).copy(),
                self._qpos.copy(), glob_out)

    def free_segment(self, segment):
        self.hal._l.r0b200_segment_free(segment[0])

    def prove_core(self, pf, two_phase=False):
        """SegmentProverImpl::prove_core (rv32im/src/prove/hal/mod.rs:143-224) from a PreflightResults, everything on
        the device in one call (r0b200_prove_segment_rv32im). Returns (seal, roots, query positions, globals).
        two_phase=True takes the protocol's own split through the Hal-level calls instead (WitnessGenerator, prove_begin
        -> mix -> accum with the host-built BigIntAccum injector -> prove_finish): same seal."""
        hal = self.hal
        if two_phase:
            wg = WitnessGenerator(hal, pf)
            try:
                glob = wg.global_.view().copy()
                h, mix = self.begin(pf.po2, wg.code, wg.data, glob)
                wg.accum(mix)
                seal, roots, qpos = self.finish(h, wg.accum_buf)
            finally:
                wg.trace.close()
            return seal, roots, qpos, glob
        st, keep = _trace_struct(pf)
        index, offsets, values = (_u32(a) for a in pf.injector)
        glob = _u32(pf.global_)
        glob_out = np.zeros(90, dtype=np.uint32)
        seal_len, nroots = C.c_size_t(0), C.c_size_t(0)
        check(hal._l.r0b200_prove_segment_rv32im(
            hal._ctx, hal.hash, C.c_uint32(pf.po2), C.byref(st), _np_ptr(glob), _np_ptr(index), C.c_size_t(index.size),
            _np_ptr(offsets), _np_ptr(values), _np_ptr(self._seal), C.c_size_t(self.seal_cap), C.byref(seal_len),
            _np_ptr(self._roots), C.c_size_t(16), C.byref(nroots), _np_ptr(self._qpos), _np_ptr(glob_out)))
        return (self._seal[:seal_len.value].copy(), self._roots[:8 * nroots.value].reshape(-1, 8).copy(),
                self._qpos.copy(), glob_out)

    def prove_uploaded(self, uploaded, glob):
        hal = self.hal
        h, _keep = uploaded
        glob = _u32(glob)
        seal_len, nroots = C.c_size_t(0), C.c_size_t(0)
        try:
            check(hal._l.r0b200_prove_uploaded(hal._ctx, hal.hash, h, _np_ptr(glob), _np_ptr(self._seal),
                                               C.c_size_t(self.seal_cap), C.byref(seal_len), _np_ptr(self._roots),
                                               C.c_size_t(16), C.byref(nroots), _np_ptr(self._qpos)))
        finally:
            hal._l.r0b200_witness_free(h)
        return (self._seal[:seal_len.value].copy(), self._roots[:8 * nroots.value].reshape(-1, 8).copy(),
                self._qpos.copy())

    def prove(self, po2, code, data, accum, glob, circuit="rv32im"):
        """returns (seal words, committed roots [k, 8], drawn query positions [50]). circuit = "rv32im"
        (prove_core, rv32im/src/prove/hal/mod.rs:171-222) or "recursion" (recursion/src/prove/mod.rs:179-224)"""
        hal = self.hal
        on_host = isinstance(data, np.ndarray)
        n = 1 << po2
        c_code, c_data, c_accum, n_glob = self.SHAPES[circuit]
        if on_host:
            code, data, accum = _u32(code), _u32(data), _u32(accum)
            assert code.size == c_code * n and data.size == c_data * n and accum.size == c_accum * n
            ptrs = [_np_ptr(code), _np_ptr(data), _np_ptr(accum)]
        else:
            assert code.size() == c_code * n and data.size() == c_data * n and accum.size() == c_accum * n
            ptrs = [code.ptr, data.ptr, accum.ptr]
        glob = _u32(glob)
        assert glob.size == n_glob
        seal_len, nroots = C.c_size_t(0), C.c_size_t(0)
        fn = hal._l.r0b200_prove_rv32im if circuit == "rv32im" else hal._l.r0b200_prove_recursion
        check(fn(hal._ctx, hal.hash, C.c_uint32(po2), ptrs[0], ptrs[1], ptrs[2],
                 C.c_int(1 if on_host else 0), _np_ptr(glob), _np_ptr(self._seal),
                 C.c_size_t(self.seal_cap), C.byref(seal_len), _np_ptr(self._roots),
                 C.c_size_t(16), C.byref(nroots), _np_ptr(self._qpos)))
        return (self._seal[:seal_len.value].copy(), self._roots[:8 * nroots.value].reshape(-1, 8).copy(),
                self._qpos.copy())
