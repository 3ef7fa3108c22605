// Expression limit states: program validation and CUDA code generation (host only).
// Program format: include/ebncuda.h "expression program".
#include "ebn_expr.h"

#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/ebncuda.h"

namespace ebn {
namespace {

struct OpShape { int pops, pushes; const char* name; };
// stack effect of each opcode
const OpShape kShape[EBN_OP_NOPS] = {
    {0, 1, "input"}, {0, 1, "param"}, {0, 1, "const"}, {2, 1, "add"}, {2, 1, "sub"}, {2, 1, "mul"},
    {2, 1, "div"},   {1, 1, "neg"},   {1, 1, "abs"},   {1, 1, "sqrt"}, {1, 1, "exp"}, {1, 1, "log"},
    {1, 1, "sin"},   {1, 1, "cos"},   {2, 1, "min"},   {2, 1, "max"}, {2, 1, "pow"}, {1, 1, "powi"},
    {2, 1, "lt"},    {3, 1, "select"}, {1, 0, "store"}};

bool is_int(double v) { return std::isfinite(v) && v == std::floor(v) && std::fabs(v) < 1e9; }

std::string fmt(const char* f, ...) {
  char buf[256];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof buf, f, ap);
  va_end(ap);
  return buf;
}

// exact literal: the bit pattern, so the generated code never depends on decimal printing
std::string lit(double v) {
  uint64_t b;
  memcpy(&b, &v, 8);
  return fmt("__longlong_as_double(0x%016llxLL)", (unsigned long long)b);
}

}  // namespace

std::string expr_check(const double* consts, int n_consts, int d, ExprInfo* info) {
  ExprInfo I;
  if (!consts || n_consts < 2) return "expression program is empty";
  if (!is_int(consts[0]) || !is_int(consts[1])) return "expression header (P, L) must be integers";
  I.n_params = (int)consts[0];
  I.n_ops = (int)consts[1];
  if (I.n_params < 0 || I.n_params > EBN_EXPR_MAX_PARAMS)
    return fmt("expression has %d runtime constants (at most %d)", I.n_params, EBN_EXPR_MAX_PARAMS);
  if (I.n_ops < 1 || I.n_ops > EBN_EXPR_MAX_OPS)
    return fmt("expression has %d instructions (1..%d)", I.n_ops, EBN_EXPR_MAX_OPS);
  if ((int64_t)n_consts != 2 + (int64_t)I.n_params + 2 * (int64_t)I.n_ops)
    return fmt("expression program: expected %lld consts for P=%d, L=%d, got %d",
               2 + (long long)I.n_params + 2 * (long long)I.n_ops, I.n_params, I.n_ops, n_consts);
  const double* code = consts + 2 + I.n_params;
  int depth = 0;
  for (int i = 0; i < I.n_ops; ++i) {
    const double opd = code[2 * i], arg = code[2 * i + 1];
    if (!is_int(opd) || opd < 0 || opd >= EBN_OP_NOPS) return fmt("instruction %d: unknown opcode %g", i, opd);
    const int op = (int)opd;
    if (depth < kShape[op].pops)
      return fmt("instruction %d (%s): needs %d operands, the stack holds %d", i, kShape[op].name,
                 kShape[op].pops, depth);
    depth += kShape[op].pushes - kShape[op].pops;
    if (depth > I.max_depth) I.max_depth = depth;
    if (depth > EBN_EXPR_MAX_DEPTH) return fmt("instruction %d: stack deeper than %d", i, EBN_EXPR_MAX_DEPTH);
    switch (op) {
      case EBN_OP_INPUT:
        if (!is_int(arg) || arg < 0 || arg >= d + I.n_store)
          return fmt("instruction %d reads column %g (only %d exist)", i, arg, d + I.n_store);
        break;
      case EBN_OP_PARAM:
        if (!is_int(arg) || arg < 0 || arg >= I.n_params)
          return fmt("instruction %d reads runtime constant %g (only %d exist)", i, arg, I.n_params);
        break;
      case EBN_OP_POWI:
        if (!is_int(arg)) return fmt("instruction %d: powi exponent %g is not an integer", i, arg);
        break;
      case EBN_OP_STORE:
        ++I.n_store;
        if (d + I.n_store > EBN_MAX_INPUTS)
          return fmt("instruction %d: %d inputs + %d stored columns exceed %d columns", i, d, I.n_store,
                     EBN_MAX_INPUTS);
        break;
      default:
        break;
    }
  }
  if (depth != 1) return fmt("expression leaves %d values on the stack (expected exactly 1: g)", depth);
  if (info) *info = I;
  return "";
}

std::string expr_code_key(const double* consts, int n_consts, int d) {
  const int P = (int)consts[0];
  std::string key = fmt("expr:d=%d:P=%d:", d, P);
  key.append((const char*)(consts + 2 + P), (size_t)(n_consts - 2 - P) * sizeof(double));
  return key;
}

// Values that depend on runtime constants and literals only are computed once per thread in
// setup() and kept in the functor's context: same operations, same rounding, just not per sample.
// A division by such a value becomes, outside strict mode, a multiplication by its reciprocal.
std::string expr_codegen(const double* consts, int n_consts, int d) {
  ExprInfo I;
  if (!expr_check(consts, n_consts, d, &I).empty()) return "";
  const double* code = consts + 2 + I.n_params;
  struct Val {
    std::string name;  // how eval() refers to it (for an invariant: how setup() refers to it)
    bool inv;          // depends on runtime constants / literals only
    bool leaf;         // a constant or literal itself: usable in eval() directly
  };
  std::string setup, eval;
  std::vector<Val> st;
  int nv = 0, ns = 0, nh = 0, col = d;
  // name of `v` as seen from eval(): invariants computed in setup() get a context slot
  auto in_eval = [&](const Val& v) {
    if (!v.inv || v.leaf) return v.name;
    setup += fmt("    c.h[%d] = ", nh) + v.name + ";\n";
    return fmt("c.h[%d]", nh++);
  };
  auto pop = [&]() {
    Val v = st.back();
    st.pop_back();
    return v;
  };
  // result of an operation on `args`; rhs(names) builds the expression from the operand names
  auto apply = [&](std::vector<Val> args, auto rhs) {
    bool inv = true;
    for (const Val& a : args) inv = inv && a.inv;
    std::vector<std::string> names;
    for (const Val& a : args) names.push_back(inv ? a.name : in_eval(a));
    if (inv) {
      const std::string v = fmt("s%d", ns++);
      setup += "    const double " + v + " = " + rhs(names) + ";\n";
      st.push_back({v, true, false});
    } else {
      const std::string v = fmt("v%d", nv++);
      eval += "    const double " + v + " = " + rhs(names) + ";\n";
      st.push_back({v, false, false});
    }
  };
  auto call2 = [](const char* f) {
    return [f](const std::vector<std::string>& n) { return std::string(f) + "(" + n[0] + ", " + n[1] + ")"; };
  };
  auto call1 = [](const char* f) {
    return [f](const std::vector<std::string>& n) { return std::string(f) + "(" + n[0] + ")"; };
  };
  for (int i = 0; i < I.n_ops; ++i) {
    const int op = (int)code[2 * i];
    const double arg = code[2 * i + 1];
    Val a, b, c3;
    if (kShape[op].pops == 3) { b = pop(); a = pop(); c3 = pop(); }
    else if (kShape[op].pops == 2) { b = pop(); a = pop(); }
    else if (kShape[op].pops == 1) { a = pop(); }
    switch (op) {
      case EBN_OP_INPUT: st.push_back({fmt("x[%d]", (int)arg), false, false}); break;
      case EBN_OP_PARAM: st.push_back({fmt("c.k[%d]", (int)arg), true, true}); break;
      case EBN_OP_CONST: st.push_back({lit(arg), true, true}); break;
      case EBN_OP_ADD: apply({a, b}, call2("Ops::add")); break;
      case EBN_OP_SUB: apply({a, b}, call2("Ops::sub")); break;
      case EBN_OP_MUL: apply({a, b}, call2("Ops::mul")); break;
      case EBN_OP_DIV:
        if (b.inv && !a.inv) {
          // strict: a / b. fast: a * (1 / b), the reciprocal taken once in setup()
          const std::string bn = b.name;
          setup += fmt("    c.h[%d] = Ops::strict ? ", nh) + bn + " : 1.0 / " + bn + ";\n";
          const std::string slot = fmt("c.h[%d]", nh++);
          const std::string an = in_eval(a);
          const std::string v = fmt("v%d", nv++);
          eval += "    const double " + v + " = Ops::strict ? Ops::div(" + an + ", " + slot + ") : " + an + " * " + slot + ";\n";
          st.push_back({v, false, false});
        } else {
          apply({a, b}, call2("Ops::div"));
        }
        break;
      case EBN_OP_NEG: apply({a}, [](const std::vector<std::string>& n) { return "-(" + n[0] + ")"; }); break;
      case EBN_OP_ABS: apply({a}, call1("fabs")); break;
      case EBN_OP_SQRT: apply({a}, call1("Ops::sqrt")); break;
      // per-sample exp / log take the lean table-driven versions outside strict mode; a constant
      // operand is evaluated in setup(), before the shared-memory tables exist: plain library call
      case EBN_OP_EXP:
        if (a.inv) apply({a}, call1("exp"));
        else apply({a}, [](const std::vector<std::string>& n) {
          return "(Ops::strict ? exp(" + n[0] + ") : exp_guarded(" + n[0] + ", *c.sm))";
        });
        break;
      case EBN_OP_LOG:
        if (a.inv) apply({a}, call1("log"));
        else apply({a}, [](const std::vector<std::string>& n) {
          return "(Ops::strict ? log(" + n[0] + ") : log_guarded(" + n[0] + ", *c.sm))";
        });
        break;
      case EBN_OP_SIN: apply({a}, call1("sin")); break;
      case EBN_OP_COS: apply({a}, call1("cos")); break;
      case EBN_OP_MIN: apply({a, b}, call2("Ops::min")); break;
      case EBN_OP_MAX: apply({a, b}, call2("Ops::max")); break;
      case EBN_OP_POW: apply({a, b}, call2("pow")); break;
      case EBN_OP_POWI: {
        const int n = (int)arg;
        apply({a}, [n](const std::vector<std::string>& v) {
          const std::string& x = v[0];
          if (n == 0) return std::string("1.0");
          if (n == 1) return x;
          if (n == 2) return "Ops::mul(" + x + ", " + x + ")";
          if (n == 3) return "Ops::mul(Ops::mul(" + x + ", " + x + "), " + x + ")";
          return "pow(" + x + ", " + fmt("%d.0", n) + ")";
        });
        break;
      }
      case EBN_OP_LT:
        apply({a, b}, [](const std::vector<std::string>& n) { return "((" + n[0] + ") < (" + n[1] + ") ? 1.0 : 0.0)"; });
        break;
      case EBN_OP_SELECT:
        apply({c3, a, b}, [](const std::vector<std::string>& n) {
          return "((" + n[0] + ") != 0.0 ? (" + n[1] + ") : (" + n[2] + "))";
        });
        break;
      case EBN_OP_STORE:
        eval += fmt("    x[%d] = ", col++) + in_eval(a) + ";\n";
        break;
    }
  }
  const std::string result = in_eval(st.back());
  std::string s;
  s += "namespace ebn {\ntemplate <class Ops>\nstruct JitExpr {\n";
  s += "  static constexpr int ID = EBN_LS_EXPRESSION;\n";
  s += "  static constexpr bool HAS_FAILS = false;\n";
  s += "  static constexpr int MIN_BLOCKS = EBN_BLOCKS_FOR_WARPS(16);\n";
  s += fmt("  static constexpr int D = %d;\n", d);
  s += fmt("  static constexpr int XCAP = %d;\n", d + I.n_store);
  s += "  static constexpr int NC = -1;\n";
  s += fmt("  struct Ctx { double k[%d]; double h[%d]; const FastMathSmem* sm; };\n", I.n_params > 0 ? I.n_params : 1,
           nh > 0 ? nh : 1);
  s += "  __device__ static void setup(Ctx& c, const double* k, int, int, const FastMathSmem* sm) {\n";
  s += "    c.sm = sm;\n";
  s += fmt("#pragma unroll\n    for (int i = 0; i < %d; ++i) c.k[i] = k[2 + i];\n", I.n_params);
  s += setup;
  s += "  }\n";
  s += "  __device__ __forceinline__ static double eval(const Ctx& c, double* x) {\n";
  s += eval;
  s += "    return " + result + ";\n  }\n};\n}  // namespace ebn\n";
  return s;
}

}  // namespace ebn
