// Runtime of the generated rv32im witness-generation code (csrc/gen/witgen_rv32im.inc, tools/gen_witgen.py): field
// value types, column-major matrix access with the reference's `checked` buffer semantics, and the preflight-trace
// externs. Follows, for behaviour, risc0/circuit/rv32im-sys/kernels/cxx/{witgen.h,buffers.h,tables.h,ffi.cpp:41-262}
// (CUDA twin kernels/cuda/{witgen.h,buffers.h,tables.h,ffi.cu:233-360}); the design is this backend's own:
//   * a bound layout is ONE 32-bit word: buffer id << 24 | index into a flat uint16 column table (the generator folds
//     every layout lookup into a constant added to that word), instead of pointers to layout objects + virtual buffers;
//   * failures (inconsistent set, read of unset value, eqz, trace mismatch) are recorded in a 4-word error block
//     (count, code, cycle, site) instead of exceptions / device asserts, so a kernel never traps the context;
//   * lookup-table increments are warp-aggregated on the device (one atomic per distinct index per warp).
// Compiles for the device (nvcc, sm_100a) and, for the CPU check of the generator in tests/, for the host (g++).
#pragma once
#include <stdint.h>

#include "fp.cuh"

#if defined(__CUDACC__)
#define WG_FN __device__
#define WG_INLINE __device__ __forceinline__
#define WG_NOINLINE __device__ __noinline__
#define WG_RT __device__ __forceinline__
#else
#define WG_FN
#define WG_INLINE inline
#define WG_NOINLINE
#define WG_RT inline
#endif

namespace r0wg {

using r0::P;

// ---- value types --------------------------------------------------------------------------------------------------
struct Val {
  uint32_t v;  // Montgomery form, canonical
  Val() = default;
  static WG_RT Val raw(uint32_t m) {
    Val r;
    r.v = m;
    return r;
  }
  static WG_RT Val from_u32(uint32_t x) { return raw(r0::fp_encode(x)); }   // normal-form integer -> field
  WG_RT uint32_t u32() const { return r0::fp_decode(v); }                    // Fp::asUInt32
};
WG_RT Val operator+(Val a, Val b) { return Val::raw(r0::fp_add(a.v, b.v)); }
WG_RT Val operator-(Val a, Val b) { return Val::raw(r0::fp_sub(a.v, b.v)); }
WG_RT Val operator*(Val a, Val b) { return Val::raw(r0::fp_mul(a.v, b.v)); }
WG_RT Val operator-(Val a) { return Val::raw(r0::fp_neg(a.v)); }

struct ExtVal {
  Val e[4];
};
WG_RT r0::FpExt to_fpext(const ExtVal& a) { return r0::FpExt{{a.e[0].v, a.e[1].v, a.e[2].v, a.e[3].v}}; }
WG_RT ExtVal from_fpext(const r0::FpExt& a) {
  return ExtVal{{Val::raw(a.c[0]), Val::raw(a.c[1]), Val::raw(a.c[2]), Val::raw(a.c[3])}};
}
WG_RT ExtVal ext_of(Val a) { return ExtVal{{a, Val::raw(0), Val::raw(0), Val::raw(0)}}; }
WG_RT ExtVal operator+(const ExtVal& a, const ExtVal& b) { return from_fpext(r0::ext_add(to_fpext(a), to_fpext(b))); }
WG_RT ExtVal operator-(const ExtVal& a, const ExtVal& b) { return from_fpext(r0::ext_sub(to_fpext(a), to_fpext(b))); }
WG_RT ExtVal operator*(const ExtVal& a, const ExtVal& b) { return from_fpext(r0::ext_mul(to_fpext(a), to_fpext(b))); }
WG_RT ExtVal operator-(const ExtVal& a) { return from_fpext(r0::ext_neg(to_fpext(a))); }
WG_RT ExtVal operator*(const ExtVal& a, Val b) { return from_fpext(r0::ext_scale(to_fpext(a), b.v)); }
WG_RT ExtVal operator*(Val a, const ExtVal& b) { return b * a; }
WG_RT ExtVal operator+(const ExtVal& a, Val b) { return a + ext_of(b); }
WG_RT ExtVal operator+(Val a, const ExtVal& b) { return ext_of(a) + b; }
WG_RT ExtVal operator-(const ExtVal& a, Val b) { return a - ext_of(b); }
WG_RT ExtVal operator-(Val a, const ExtVal& b) { return ext_of(a) - b; }

template <typename T, int N>
struct Arr {
  T v[N];
  WG_RT T& operator[](uint32_t i) { return v[i]; }
  WG_RT const T& operator[](uint32_t i) const { return v[i]; }
};

// built-in field operations of the step functions (witgen.h:57-93)
WG_RT bool nz(Val x) { return x.v != 0; }
WG_RT uint32_t to_size_t(Val x) { return x.u32(); }
WG_RT Val isz(Val x) { return x.v == 0 ? Val::raw(r0::MONT_ONE) : Val::raw(0); }
WG_RT Val neg_0(Val x) { return -x; }
WG_RT Val inv_0(Val x) { return Val::raw(r0::fp_inv(x.v)); }
WG_RT ExtVal inv_0(const ExtVal& x) { return from_fpext(r0::ext_inv(to_fpext(x))); }
WG_RT Val bitAnd(Val a, Val b) { return Val::from_u32(a.u32() & b.u32()); }
WG_RT Val mod(Val a, Val b) { return Val::from_u32(a.u32() % b.u32()); }
WG_RT Val inRange(Val low, Val mid, Val high) {
  const uint32_t m = mid.u32();
  return (low.u32() <= m && m < high.u32()) ? Val::raw(r0::MONT_ONE) : Val::raw(0);
}

// ---- preflight trace (rv32im-sys/src/lib.rs:21-84 RawPreflightCycle 36 B, RawMemoryTransaction 20 B) ---------------
struct PreflightCycle {
  uint32_t state, pc;
  uint8_t major, minor, machine_mode, padding;
  uint32_t user_cycle, txn_idx, paging_idx, bigint_idx;
  uint32_t diff_count[2];
};
struct MemoryTxn {
  uint32_t addr, cycle, word, prev_cycle, prev_word;
};
static_assert(sizeof(PreflightCycle) == 36 && sizeof(MemoryTxn) == 20, "preflight trace layout");

enum { BUF_DATA = 0, BUF_ACCUM = 1, BUF_GLOBAL = 2, BUF_MIX = 3 };
enum {  // error codes recorded in WShared::err[1]
  WG_ERR_INCONSISTENT_SET = 1,
  WG_ERR_READ_UNSET = 2,
  WG_ERR_EQZ = 3,
  WG_ERR_TXN_CYCLE = 4,
  WG_ERR_TXN_ADDR = 5,
  WG_ERR_LOOKUP_TABLE = 6,
  WG_ERR_LOOKUP_RANGE = 7,
  WG_ERR_UNREACHABLE = 8,
};

struct WBuf {
  uint32_t* buf;
  uint32_t rows, cols;
  uint32_t checked;    // Buffer::checked (buffers.h)
  uint32_t zero_back;  // MutableBufObj::zeroBack (witgen.h:113-118): back-loads of columns above it read as 0
};

// launch-wide state
struct WShared {
  WBuf bufs[4];
  const PreflightCycle* cycles;
  const MemoryTxn* txns;
  const uint8_t* bigint_bytes;
  const uint16_t* layout;  // flat column table (R0_WG_LAYOUT_DATA)
  uint32_t* table_u8;      // LookupTables (tables.h): 2^8 and 2^16 counters
  uint32_t* table_u16;
  uint32_t* err;           // [count, code, cycle, site / detail]
  uint32_t txns_len, bigint_len;
};

// per-cycle state (ExecContext, witgen.h:95-101)
struct WCtx {
  const WShared* s;
  uint32_t cycle;
  uint32_t txn_cursor;  // cycles[cycle].txn_idx, advanced by getMemoryTxn (the reference increments it in the trace)
};

typedef uint32_t BL;  // bound layout: buffer id << 24 | index into the column table
WG_RT BL bind_layout(uint32_t base, uint32_t buf) { return (buf << 24) | base; }

WG_RT void fail(WCtx& ctx, uint32_t code, uint32_t detail) {
#if defined(__CUDA_ARCH__)
  if (atomicAdd(ctx.s->err, 1u) == 0u) {
#else
  if (__atomic_fetch_add(ctx.s->err, 1u, __ATOMIC_RELAXED) == 0u) {
#endif
    ctx.s->err[1] = code;
    ctx.s->err[2] = ctx.cycle;
    ctx.s->err[3] = detail;
  }
}

// Buffer::get / MutableBufObj::load / GlobalBufObj::load (buffers.h:46-54, witgen.h:113-148)
WG_RT Val ld(WCtx& ctx, BL bl, uint32_t back) {
  const uint32_t col = ctx.s->layout[bl & 0xffffffu];
  const uint32_t kind = bl >> 24;
  const WBuf& b = ctx.s->bufs[kind];
  uint32_t w;
  if (kind >= BUF_GLOBAL) {
    w = b.buf[col];
  } else {
    if (b.zero_back && col > b.zero_back && back > 0) return Val::raw(0);
    const uint32_t row = (ctx.cycle + b.rows - back) & (b.rows - 1);   // rows is a power of two
    w = b.buf[(size_t)col * b.rows + row];
  }
  if (w == r0::FP_INVALID && b.checked) fail(ctx, WG_ERR_READ_UNSET, (kind << 24) | col);
  return Val::raw(w);
}
WG_RT ExtVal ldext(WCtx& ctx, BL bl, uint32_t back) {
  // loadExt: the four columns col .. col + 3 of the register's first column (witgen.h:201-207)
  const uint32_t col = ctx.s->layout[bl & 0xffffffu];
  const uint32_t kind = bl >> 24;
  const WBuf& b = ctx.s->bufs[kind];
  ExtVal r;
  for (uint32_t i = 0; i < 4; i++) {
    uint32_t w;
    if (kind >= BUF_GLOBAL) {
      w = b.buf[col + i];
    } else if (b.zero_back && col + i > b.zero_back && back > 0) {
      w = 0;
    } else {
      const uint32_t row = (ctx.cycle + b.rows - back) & (b.rows - 1);
      w = b.buf[(size_t)(col + i) * b.rows + row];
    }
    if (w == r0::FP_INVALID && b.checked) fail(ctx, WG_ERR_READ_UNSET, (kind << 24) | (col + i));
    r.e[i] = Val::raw(w);
  }
  return r;
}
// Buffer::set (buffers.h:30-44)
WG_RT void st_col(WCtx& ctx, uint32_t kind, uint32_t col, Val val) {
  const WBuf& b = ctx.s->bufs[kind];
  uint32_t* p = kind >= BUF_GLOBAL ? b.buf + col : b.buf + (size_t)col * b.rows + ctx.cycle;
  const uint32_t cur = *p;
  if (cur != r0::FP_INVALID && cur != val.v && b.checked) fail(ctx, WG_ERR_INCONSISTENT_SET, (kind << 24) | col);
  *p = val.v;
}
WG_RT void st(WCtx& ctx, BL bl, Val val) { st_col(ctx, bl >> 24, ctx.s->layout[bl & 0xffffffu], val); }
WG_RT void stext(WCtx& ctx, BL bl, const ExtVal& val) {
  const uint32_t col = ctx.s->layout[bl & 0xffffffu];
  for (uint32_t i = 0; i < 4; i++) st_col(ctx, bl >> 24, col + i, val.e[i]);
}

WG_RT void eqz(WCtx& ctx, Val a, uint32_t site) {
  if (a.v != 0) fail(ctx, WG_ERR_EQZ, site);
}
WG_RT void eqz(WCtx& ctx, const ExtVal& a, uint32_t site) {
  if ((a.e[0].v | a.e[1].v | a.e[2].v | a.e[3].v) != 0) fail(ctx, WG_ERR_EQZ, site);
}
WG_RT void unreachable(WCtx& ctx) { fail(ctx, WG_ERR_UNREACHABLE, 0); }

// ---- externs (ffi.cpp:85-262) --------------------------------------------------------------------------------------
WG_RT Arr<Val, 5> ext_getMemoryTxn(WCtx& ctx, Val addr_elem) {
  const uint32_t addr = addr_elem.u32();
  const uint32_t idx = ctx.txn_cursor++;
  Arr<Val, 5> r;
  if (idx >= ctx.s->txns_len) {
    fail(ctx, WG_ERR_TXN_CYCLE, idx);
    for (int i = 0; i < 5; i++) r[i] = Val::raw(0);
    return r;
  }
  const MemoryTxn t = ctx.s->txns[idx];
  if (t.cycle / 2 != ctx.cycle) fail(ctx, WG_ERR_TXN_CYCLE, idx);
  if (t.addr != addr) fail(ctx, WG_ERR_TXN_ADDR, idx);
  r[0] = Val::from_u32(t.prev_cycle);
  r[1] = Val::from_u32(t.prev_word & 0xffff);
  r[2] = Val::from_u32(t.prev_word >> 16);
  r[3] = Val::from_u32(t.word & 0xffff);
  r[4] = Val::from_u32(t.word >> 16);
  return r;
}

WG_RT void table_inc(uint32_t* table, uint32_t index) {
#if defined(__CUDA_ARCH__)
  // warp-aggregated: lanes hitting the same counter elect one leader that adds the group's population
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, index);
  if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(table + index, (uint32_t)__popc(peers));
#else
  __atomic_fetch_add(table + index, 1u, __ATOMIC_RELAXED);
#endif
}
// LookupTables::lookupDelta (tables.h:30-52): the count argument is ignored by the reference, every call adds one
WG_RT void ext_lookupDelta(WCtx& ctx, Val table, Val index, Val /*count*/) {
  const uint32_t t = table.u32(), i = index.u32();
  if (t == 0) return;
  if (t != 8 && t != 16) {
    fail(ctx, WG_ERR_LOOKUP_TABLE, t);
    return;
  }
  if (i >= (1u << t)) {
    fail(ctx, WG_ERR_LOOKUP_RANGE, i);
    return;
  }
  table_inc(t == 8 ? ctx.s->table_u8 : ctx.s->table_u16, i);
}
WG_RT Val ext_lookupCurrent(WCtx& ctx, Val table, Val index) {
  const uint32_t t = table.u32(), i = index.u32();
  if (t != 8 && t != 16) {
    fail(ctx, WG_ERR_LOOKUP_TABLE, t);
    return Val::raw(0);
  }
  if (i >= (1u << t)) {
    fail(ctx, WG_ERR_LOOKUP_RANGE, i);
    return Val::raw(0);
  }
  return Val::from_u32((t == 8 ? ctx.s->table_u8 : ctx.s->table_u16)[i]);
}
WG_RT void ext_memoryDelta(WCtx&, Val, Val, Val, Val, Val) {}
WG_RT Val ext_getDiffCount(WCtx& ctx, Val cycle) {
  const uint32_t c = cycle.u32();
  return Val::from_u32(ctx.s->cycles[c / 2].diff_count[c % 2]);
}
WG_RT Val ext_isFirstCycle_0(WCtx& ctx) { return ctx.cycle == 0 ? Val::raw(r0::MONT_ONE) : Val::raw(0); }

// divide_rv32im (ffi.cpp:53-83)
WG_RT Arr<Val, 4> ext_divide(WCtx&, Val numer_low, Val numer_high, Val denom_low, Val denom_high, Val sign_type) {
  uint32_t numer = numer_low.u32() | (numer_high.u32() << 16);
  uint32_t denom = denom_low.u32() | (denom_high.u32() << 16);
  const uint32_t sign = sign_type.u32();
  const uint32_t ones_comp = sign == 2;
  const bool neg_numer = sign && (int32_t)numer < 0;
  const bool neg_denom = sign == 1 && (int32_t)denom < 0;
  if (neg_numer) numer = 0u - numer - ones_comp;
  if (neg_denom) denom = 0u - denom - ones_comp;
  uint32_t quot, rem;
  if (denom == 0) {
    quot = 0xffffffffu;
    rem = numer;
  } else {
    quot = numer / denom;
    rem = numer % denom;
  }
  const uint32_t quot_neg = (uint32_t)(neg_numer ^ neg_denom) - ((denom == 0) * (uint32_t)neg_numer);
  if (quot_neg) quot = 0u - quot - ones_comp;
  if (neg_numer) rem = 0u - rem - ones_comp;
  Arr<Val, 4> r;
  r[0] = Val::from_u32(quot & 0xffff);
  r[1] = Val::from_u32(quot >> 16);
  r[2] = Val::from_u32(rem & 0xffff);
  r[3] = Val::from_u32(rem >> 16);
  return r;
}
WG_RT Arr<Val, 2> ext_getMajorMinor(WCtx& ctx) {
  const PreflightCycle& c = ctx.s->cycles[ctx.cycle];
  Arr<Val, 2> r;
  r[0] = Val::from_u32(c.major);
  r[1] = Val::from_u32(c.minor);
  return r;
}
// both read the word of the NEXT unconsumed transaction of this cycle (ffi.cpp:217-229)
WG_RT Val ext_hostReadPrepare(WCtx& ctx, Val, Val) { return Val::from_u32(ctx.s->txns[ctx.txn_cursor].word); }
WG_RT Val ext_hostWrite(WCtx& ctx, Val, Val, Val, Val) { return Val::from_u32(ctx.s->txns[ctx.txn_cursor].word); }
WG_RT Arr<Val, 2> ext_nextPagingIdx(WCtx& ctx) {
  const PreflightCycle& c = ctx.s->cycles[ctx.cycle];
  Arr<Val, 2> r;
  r[0] = Val::from_u32(c.paging_idx);
  r[1] = Val::from_u32(c.machine_mode);
  return r;
}
WG_RT Arr<Val, 16> ext_bigIntExtern(WCtx& ctx) {
  const uint32_t base = ctx.s->cycles[ctx.cycle].bigint_idx;
  Arr<Val, 16> r;
  for (uint32_t i = 0; i < 16; i++) r[i] = Val::from_u32(base + i < ctx.s->bigint_len ? ctx.s->bigint_bytes[base + i] : 0u);
  return r;
}
WG_RT void ext_print(WCtx&, Val) {}

}  // namespace r0wg
