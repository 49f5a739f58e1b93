// ORACLE — TEST INFRASTRUCTURE ONLY. C-linkage shim around the reference's OWN generated recursion constraint code.
// Nothing is copied: the reference source is compiled where it lies (the Makefile passes its directory with -I), and
// this file only adds an extern "C" entry point in the shape the rv32im reference already exports
// (risc0_circuit_rv32im_cpu_poly_fp, rv32im-sys/kernels/cxx/eval_check.cpp:30-38); the recursion tree only exports
// the whole-domain risc0_circuit_recursion_cpu_eval_check (recursion-sys/kernels/cxx/ffi.cpp:219-246), which drags in
// the witgen contexts.
#include "poly_fp.cpp"  // risc0/circuit/recursion-sys/kernels/cxx/poly_fp.cpp, found through -I

#include <cstring>
#include <exception>

extern "C" const char* r0ref_recursion_poly_fp(size_t cycle, size_t steps, risc0::FpExt* poly_mix, risc0::Fp** args,
                                               risc0::FpExt* result) {
  try {
    *result = risc0::circuit::recursion::poly_fp(cycle, steps, poly_mix, args);
  } catch (const std::exception& e) {
    return strdup(e.what());
  }
  return nullptr;
}
