// TEST INFRASTRUCTURE ONLY. Host (g++) build of the GENERATED witness-generation code (csrc/gen/witgen_rv32im.inc +
// csrc/witgen_rt.cuh), so that tools/gen_witgen.py and the circuit IR can be checked on the CPU, in the build container,
// against the reference's own compiled witgen (oracle/_ref/librv32im_witgen_ref.so) before the same text is compiled
// for the device. Never loaded by the product (the product's witgen is csrc/witgen.cu and needs a GPU).
#include <string.h>

#include <vector>

#include "../risc0_b200/csrc/witgen_rt.cuh"

namespace r0wg {
#include "../risc0_b200/csrc/gen/witgen_rv32im.inc"
static const uint16_t kLayoutTable[] = R0_WG_LAYOUT_DATA;
}  // namespace r0wg

using namespace r0wg;

static void setup(WShared& s, const PreflightCycle* cycles, const MemoryTxn* txns, uint32_t ntxns, const uint8_t* bigint,
                  uint32_t nbig, std::vector<uint32_t>& tables, uint32_t* err) {
  memset(&s, 0, sizeof(s));
  s.cycles = cycles;
  s.txns = txns;
  s.txns_len = ntxns;
  s.bigint_bytes = bigint;
  s.bigint_len = nbig;
  s.layout = kLayoutTable;
  tables.assign(256 + 65536, 0);
  s.table_u8 = tables.data();
  s.table_u16 = tables.data() + 256;
  s.err = err;
}

extern "C" {

// risc0_circuit_rv32im_cpu_witgen (ffi.cpp:275-314), parallel mode: cycles [0, split) then [split, ncycles)
void wg_host_witgen(const void* cycles, uint32_t ncycles, const void* txns, uint32_t ntxns, const uint8_t* bigint,
                    uint32_t nbig, uint32_t split, uint32_t* global, uint32_t* data, uint32_t rows, uint32_t* err) {
  WShared s;
  std::vector<uint32_t> tables;
  setup(s, (const PreflightCycle*)cycles, (const MemoryTxn*)txns, ntxns, bigint, nbig, tables, err);
  s.bufs[BUF_DATA] = WBuf{data, rows, R0_WG_KREGCOUNTDATA, 1, 0};
  s.bufs[BUF_GLOBAL] = WBuf{global, 1, R0_WG_KREGCOUNTGLOBAL, 1, 0};
  for (int phase = 0; phase < 2; phase++) {
    const uint32_t b = phase ? split : 0, e = phase ? ncycles : split;
#pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t c = b; c < e; c++) {
      WCtx ctx{&s, c, s.cycles[c].txn_idx};
      step_Top(ctx, BUF_DATA, BUF_GLOBAL);
    }
  }
}

// risc0_circuit_rv32im_cpu_accum (ffi.cpp:316-365)
void wg_host_accum(const void* cycles, uint32_t ncycles, const void* txns, uint32_t ntxns, const uint8_t* bigint,
                   uint32_t nbig, uint32_t* data, uint32_t* accum, uint32_t* global, uint32_t* mix, uint32_t rows,
                   uint32_t* err) {
  WShared s;
  std::vector<uint32_t> tables;
  setup(s, (const PreflightCycle*)cycles, (const MemoryTxn*)txns, ntxns, bigint, nbig, tables, err);
  const uint32_t split = R0_WG_USER_ACCUM_SPLIT;
  const uint32_t acols = R0_WG_KREGCOUNTACCUM;
  s.bufs[BUF_DATA] = WBuf{data, rows, R0_WG_KREGCOUNTDATA, 1, 0};
  s.bufs[BUF_ACCUM] = WBuf{accum, rows, acols, 1, split};
  s.bufs[BUF_GLOBAL] = WBuf{global, 1, R0_WG_KREGCOUNTGLOBAL, 1, 0};
  s.bufs[BUF_MIX] = WBuf{mix, 1, R0_WG_KREGCOUNTMIX, 1, 0};
#pragma omp parallel for schedule(dynamic, 256)
  for (uint32_t c = 0; c < ncycles; c++) {
    WCtx ctx{&s, c, s.cycles[c].txn_idx};
    step_TopAccum(ctx, BUF_ACCUM, BUF_DATA, BUF_GLOBAL, BUF_MIX);
  }
  for (uint32_t j = 0; j < 4; j++) {
    uint32_t* col = accum + (size_t)(acols - 4 + j) * rows;
    for (uint32_t r = 1; r < ncycles; r++) col[r] = r0::fp_add(col[r], col[r - 1]);
  }
  const uint32_t machine_cols = (acols - split) / 4;
  std::vector<uint32_t> last(4 * (size_t)ncycles);
  for (uint32_t k = 0; k < 4; k++) memcpy(&last[(size_t)k * ncycles], accum + (size_t)(acols - 4 + k) * rows, 4 * (size_t)ncycles);
  for (uint32_t row = 0; row < ncycles; row++) {
    const uint32_t back1 = (row + ncycles - 1) % ncycles;
    for (uint32_t j = 0; j + 1 < machine_cols; j++)
      for (uint32_t k = 0; k < 4; k++) {
        uint32_t* p = accum + (size_t)(split + j * 4 + k) * rows + row;
        *p = r0::fp_add(*p, last[(size_t)k * ncycles + back1]);
      }
  }
}

uint32_t wg_num_sites() { return 0; }
}
