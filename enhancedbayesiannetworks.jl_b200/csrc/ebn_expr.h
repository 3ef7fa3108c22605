// Expression limit states (EBN_LS_EXPRESSION): validation of the stack program a caller passes in
// ebn_job_t.consts and its translation into a device functor, which ebn_jit.cu compiles at run time.
// Host code only; no CUDA calls.
//
// This is the route for the reference's arbitrary `Model(df -> ..., :name)` closures and
// `performance` functions (src/nodes/functionalnodes.jl:27-52): the wrapper translates the Julia
// expression of each model into one STORE-terminated run of the program, in Julia's evaluation
// order, so strict mode reproduces `g` operation by operation.
#pragma once
#include <string>
#include <vector>

namespace ebn {

struct ExprInfo {
  int n_params = 0;   // P: runtime constants k[0..P)
  int n_ops = 0;      // L
  int n_store = 0;    // T: model columns the program appends after the d inputs
  int max_depth = 0;  // evaluation stack depth
};

// Checks the program for `d` inputs. Returns an empty string when it is valid, else the reason.
std::string expr_check(const double* consts, int n_consts, int d, ExprInfo* info);

// The part of the program that determines the generated code (everything but the values of the
// runtime constants), as bytes, for the kernel cache key.
std::string expr_code_key(const double* consts, int n_consts, int d);

// CUDA source of `template <class Ops> struct JitExpr` implementing the program (valid programs only).
std::string expr_codegen(const double* consts, int n_consts, int d);

}  // namespace ebn
