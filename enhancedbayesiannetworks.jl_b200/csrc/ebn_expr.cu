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

// ---------------------------------------------------------------------------------------------------
// Sign-only failure predicate for fast mode.
//
// The failure count needs `g < 0` (NaN counts as safe, like `sum(g .< 0)` in the reference), not g. The
// hand-written vehicle functor exploits that (VehicleSuspension::fails); this does the same for formulas:
//   min(a, b) < 0        <=>  (a < 0 or b < 0) and neither is NaN            (Julia's minimum propagates NaN)
//   a - b < 0            <=>  a < b
//   c sqrt(z) < k        <=>  k/c > 0 and 0 <= z < (k/c)^2                   (c a positive literal, k constant)
//   k < c sqrt(z)        <=>  z >= 0 and (k/c < 0 or z > (k/c)^2)
//   P/Q compared with 0  <=>  the signs of P and Q, where a value built from + - * / over other values is
//                             carried as a fraction (numerator, denominator) and never divided out
//   sqrt(z), log(z) NaN  <=>  not (z >= 0)
// Everything else falls back to the value eval() computes. Comparisons of fractions use the SIGNS of
// numerator and denominator, not their product (no spurious under/overflow). A divisor that is exactly zero or
// infinite (a Parameter state of 0, a self-cancelling formula, a literal -0.0 — never a continuous random input)
// makes IEEE arithmetic produce infinities, signed zeros and, from them, NaNs that no sign rule sees: every
// divisor the rewriting looked through is tested, and a sample with such a one takes eval()'s own arithmetic
// instead (a cold block; the condition is the same for all samples of a scenario in practice). What is left different from `eval(x) < 0`:
// roundings (ties), and products of numerators / denominators that overflow where the quotient would not
// (magnitudes beyond 1e150). Strict mode never uses the predicate.
// ---------------------------------------------------------------------------------------------------
struct Node {
  int op;
  double arg;            // literal value / column / powi exponent
  int a = -1, b = -1;    // operand nodes (select and lt are opaque to the rewriter)
  bool inv = false, leaf = false;
  std::string name;      // how eval() refers to the value (an invariant that is not a leaf: how setup() does)
  std::string slot;      // such an invariant's name inside eval()/fails(), once one was asked for
};

// Truth values are opaque ints. Written with bools (bitwise or logical operators alike), the predicate was
// miscompiled twice — found by re-seeding the fuzz test (tests/test_gpu_adversarial.py), never on the host:
//   * nvcc 12.9 folds  a = (k != k), b = (k != k), c = b; ((p | q) & !c) ... ((r | s) & !(c | a))  down to one
//     comparison, the p | q terms vanish, even at -O0 (a minimal kernel reproduces it);
//   * NVRTC 12.9, inside the fused kernel only, drops the `|| p24` term of  ((p29 || p24) && !(p3 || p8))  nested
//     under another `&& !(...)` with repeated NaN flags (the same text in a small kernel compiles correctly).
// Both are boolean re-derivations by the optimiser. Every temporary therefore goes through an empty asm that makes
// it opaque: the logic is evaluated exactly as written (one SETP + SEL per comparison, LOP3 for the rest).
struct FailsGen {
  std::vector<Node>& nodes;
  std::string& setup;
  int& nh;
  std::string code;   // statements appended after eval()'s body
  int nt = 0;
  std::vector<int> div_memo;
  std::vector<std::string> neg_memo, nan_memo;

  FailsGen(std::vector<Node>& n, std::string& s, int& h) : nodes(n), setup(s), nh(h) {
    div_memo.assign(n.size(), -1);
    neg_memo.assign(n.size(), "");
    nan_memo.assign(n.size(), "");
  }
  // value of a node inside fails()
  std::string val(int id) {
    Node& n = nodes[id];
    if (!n.inv || n.leaf) return n.name;
    if (n.slot.empty()) {
      setup += fmt("    c.h[%d] = ", nh) + n.name + ";\n";
      n.slot = fmt("c.h[%d]", nh++);
    }
    return n.slot;
  }
  // an invariant computed in setup() from an expression over setup-scope names
  std::string inv_slot(const std::string& expr) {
    setup += fmt("    c.h[%d] = ", nh) + expr + ";\n";
    return fmt("c.h[%d]", nh++);
  }
  std::string setup_name(int id) { return nodes[id].name; }   // invariants only
  std::string dtmp(const std::string& e) {
    const std::string v = fmt("r%d", nt++);
    code += "    const double " + v + " = " + e + ";\n";
    return v;
  }
  // A truth value is an int (0 / 1) that the optimiser cannot see through: the empty asm makes every temporary
  // opaque, so the compiler evaluates the logic as written instead of re-deriving it with boolean algebra (see the
  // note above FailsGen). cmp(): one comparison; btmp(): a combination of temporaries.
  std::string cmp(const std::string& e) {
    const std::string v = fmt("p%d", nt++);
    code += "    int " + v + " = (" + e + ") ? 1 : 0;\n    asm volatile(\"\" : \"+r\"(" + v + "));\n";
    return v;
  }
  std::string btmp(const std::string& e) {
    if (e == "0" || e == "1") return e;
    if (e.size() > 1 && e[0] == 'p' && e.find_first_not_of("0123456789", 1) == std::string::npos) return e;   // already a temporary
    const std::string v = fmt("p%d", nt++);
    code += "    int " + v + " = " + e + ";\n    asm volatile(\"\" : \"+r\"(" + v + "));\n";
    return v;
  }
  static std::string b_or(const std::string& x, const std::string& y) {
    if (x == "1" || y == "1") return "1";
    if (x == "0") return y;
    if (y == "0") return x;
    return "(" + x + " | " + y + ")";
  }
  static std::string b_and(const std::string& x, const std::string& y) {
    if (x == "0" || y == "0") return "0";
    if (x == "1") return y;
    if (y == "1") return x;
    return "(" + x + " & " + y + ")";
  }
  static std::string b_not(const std::string& x) {
    if (x == "0") return "1";
    if (x == "1") return "0";
    return "(" + x + " ^ 1)";
  }
  bool arith(int op) const {
    return op == EBN_OP_ADD || op == EBN_OP_SUB || op == EBN_OP_MUL || op == EBN_OP_DIV || op == EBN_OP_NEG;
  }
  // does a per-sample division sit under this node, reachable through + - * / neg ^2 only?
  bool has_div(int id) {
    if (div_memo[id] >= 0) return div_memo[id] != 0;
    const Node& n = nodes[id];
    bool r = false;
    if (!n.inv) {
      if (n.op == EBN_OP_DIV) r = !nodes[n.b].inv || has_div(n.a);
      else if (arith(n.op)) r = has_div(n.a) || (n.b >= 0 && has_div(n.b));
      else if (n.op == EBN_OP_POWI && (int)n.arg == 2) r = has_div(n.a);
    }
    div_memo[id] = r ? 1 : 0;
    return r;
  }
  struct Frac { std::string num, den; };   // den empty: 1
  std::string mul(const std::string& x, const std::string& y) {
    if (x.empty()) return y;
    if (y.empty()) return x;
    return dtmp(x + " * " + y);
  }
  std::vector<Frac> frac_memo;
  std::vector<char> frac_done;
  std::vector<std::string> zero_tests;   // "this divisor is exactly zero", one per division looked through
  Frac frac(int id) {
    if (frac_done.empty()) {
      frac_memo.resize(nodes.size());
      frac_done.assign(nodes.size(), 0);
    }
    if (!frac_done[id]) {
      frac_memo[id] = frac_build(id);
      frac_done[id] = 1;
    }
    return frac_memo[id];
  }
  Frac frac_build(int id) {
    const Node& n = nodes[id];
    if (!has_div(id)) return {val(id), ""};
    switch (n.op) {
      case EBN_OP_DIV: {
        const Frac a = frac(n.a), b = frac(n.b);
        // the divisor is b.num / b.den: zero or not finite (a literal -0.0 under another division does that) -> cold block
        zero_tests.push_back(b_not(b_and(cmp("fabs(" + b.num + ") > 0.0"), cmp("fabs(" + b.num + ") < __longlong_as_double(0x7ff0000000000000LL)"))));
        return {mul(a.num, b.den), mul(a.den, b.num)};
      }
      case EBN_OP_MUL: {
        const Frac a = frac(n.a), b = frac(n.b);
        return {mul(a.num, b.num), mul(a.den, b.den)};
      }
      case EBN_OP_NEG: {
        const Frac a = frac(n.a);
        return {dtmp("-" + a.num), a.den};
      }
      case EBN_OP_POWI: {
        const Frac a = frac(n.a);
        return {mul(a.num, a.num), mul(a.den, a.den)};
      }
      default: {   // add, sub
        const Frac a = frac(n.a), b = frac(n.b);
        const char* pm = n.op == EBN_OP_ADD ? " + " : " - ";
        if (a.den.empty()) return {dtmp(mul(a.num, b.den) + pm + b.num), b.den};
        if (b.den.empty()) return {dtmp(a.num + pm + mul(b.num, a.den)), a.den};
        return {dtmp(mul(a.num, b.den) + pm + mul(b.num, a.den)), mul(a.den, b.den)};
      }
    }
  }
  // sign tests of a fraction (false when either part is NaN, or the denominator is zero)
  std::string signs(const Frac& f, const char* n_pos_side, const char* n_neg_side) {
    // (N <op1> 0 and D > 0) or (N <op2> 0 and D < 0)
    const std::string a = cmp(f.num + " " + n_pos_side + " 0.0"), b = cmp(f.den + " > 0.0");
    const std::string c = cmp(f.num + " " + n_neg_side + " 0.0"), d = cmp(f.den + " < 0.0");
    return btmp(b_or(b_and(a, b), b_and(c, d)));
  }
  std::string lt0(const Frac& f) { return f.den.empty() ? cmp(f.num + " < 0.0") : signs(f, "<", ">"); }
  std::string gt0(const Frac& f) { return f.den.empty() ? cmp(f.num + " > 0.0") : signs(f, ">", "<"); }
  std::string ge0(const Frac& f) { return f.den.empty() ? cmp(f.num + " >= 0.0") : signs(f, ">=", "<="); }
  // node minus a name, as a fraction
  Frac minus(int id, const std::string& name) {
    const Frac z = frac(id);
    if (z.den.empty()) return {dtmp(z.num + " - " + name), ""};
    return {dtmp(z.num + " - " + mul(name, z.den)), z.den};
  }
  // c * sqrt(z) with c a positive literal (or absent): true and (z, c) on a match
  bool scaled_sqrt(int id, int* z, double* c) {
    const Node& n = nodes[id];
    if (n.op == EBN_OP_SQRT) { *z = n.a, *c = 1.0; return true; }
    if (n.op != EBN_OP_MUL) return false;
    for (int swap = 0; swap < 2; ++swap) {
      const Node& k = nodes[swap ? n.b : n.a];
      const Node& r = nodes[swap ? n.a : n.b];
      if (k.op == EBN_OP_CONST && k.arg > 0.0 && std::isfinite(k.arg) && r.op == EBN_OP_SQRT) {
        *z = r.a, *c = k.arg;
        return true;
      }
    }
    return false;
  }
  // value(a) < value(b), false if either is NaN
  std::string less(int a, int b) {
    int z;
    double c;
    if (nodes[b].inv && !nodes[a].inv && scaled_sqrt(a, &z, &c)) {          // c sqrt(z) < k
      const std::string t = inv_slot(setup_name(b) + " / " + lit(c));
      const std::string t2 = inv_slot(t + " * " + t);
      return btmp(b_and(b_and(cmp(t + " > 0.0"), ge0(frac(z))), lt0(minus(z, t2))));
    }
    if (nodes[a].inv && !nodes[b].inv && scaled_sqrt(b, &z, &c)) {          // k < c sqrt(z)
      const std::string t = inv_slot(setup_name(a) + " / " + lit(c));
      const std::string t2 = inv_slot(t + " * " + t);
      return btmp(b_and(ge0(frac(z)), b_or(cmp(t + " < 0.0"), gt0(minus(z, t2)))));
    }
    if (has_div(a) || has_div(b)) {
      const Frac x = frac(a), y = frac(b);
      Frac diff;
      if (x.den.empty()) diff = {dtmp(mul(x.num, y.den) + " - " + y.num), y.den};
      else if (y.den.empty()) diff = {dtmp(x.num + " - " + mul(y.num, x.den)), x.den};
      else diff = {dtmp(mul(x.num, y.den) + " - " + mul(y.num, x.den)), mul(x.den, y.den)};
      return lt0(diff);
    }
    return cmp(val(a) + " < " + val(b));
  }
  // is the value NaN? ("false" when it cannot be, up to the caveats in the header comment)
  std::string nanp(int id) {
    if (!nan_memo[id].empty()) return nan_memo[id];
    const Node& n = nodes[id];
    std::string r;
    if (n.op == EBN_OP_INPUT || n.op == EBN_OP_LT || (n.op == EBN_OP_CONST && n.arg == n.arg)) r = "0";
    else if (n.inv) r = cmp("isnan(" + val(id) + ")");   // a constant sub-expression may well be NaN
    else if (n.op == EBN_OP_SQRT || n.op == EBN_OP_LOG) r = btmp(b_not(ge0(frac(n.a))));
    else if (arith(n.op) || n.op == EBN_OP_MIN || n.op == EBN_OP_MAX)
      r = btmp(b_or(nanp(n.a), n.b >= 0 ? nanp(n.b) : "0"));
    else if (n.op == EBN_OP_ABS || n.op == EBN_OP_EXP || (n.op == EBN_OP_POWI && (int)n.arg >= 0)) r = nanp(n.a);
    else r = cmp("isnan(" + val(id) + ")");
    return nan_memo[id] = r;
  }
  // value < 0, false if it is NaN
  std::string negp(int id) {
    if (!neg_memo[id].empty()) return neg_memo[id];
    const Node& n = nodes[id];
    std::string r;
    if (n.inv) r = cmp(val(id) + " < 0.0");
    else if (n.op == EBN_OP_MIN)
      r = btmp(b_and(b_or(negp(n.a), negp(n.b)), b_not(b_or(nanp(n.a), nanp(n.b)))));
    else if (n.op == EBN_OP_MAX) r = btmp(b_and(negp(n.a), negp(n.b)));
    else if (n.op == EBN_OP_SUB) r = less(n.a, n.b);
    else if (has_div(id)) r = lt0(frac(id));
    else r = cmp(val(id) + " < 0.0");
    return neg_memo[id] = r;
  }
};

// eval_body / eval_result: eval()'s statements and the name of g, pasted into the cold block.
std::string expr_fails_codegen(std::vector<Node>& nodes, int root, int /*d*/, std::string& setup, int& nh,
                               const std::string& eval_body, const std::string& eval_result) {
  FailsGen g(nodes, setup, nh);
  const std::string pred = g.negp(root);
  std::string s = g.code;
  if (!g.zero_tests.empty()) {
    std::string z;
    for (const std::string& t : g.zero_tests) z += (z.empty() ? "" : " | ") + t;
    s += "    if ((" + z + ") != 0) {   // a divisor is zero or not finite: IEEE infinities, signed zeros and NaNs as eval() produces them\n";
    s += eval_body;
    s += "    return " + eval_result + " < 0.0;\n    }\n";
  }
  return s + "    return (" + pred + ") != 0;\n";
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
    int node = -1;     // this value in the expression graph (below)
  };
  std::vector<Node> nodes;     // the expression graph behind the SSA names (read by expr_fails_codegen)
  std::vector<int> col_node;   // node of every stored column
  std::string setup, eval;
  std::vector<Val> st;
  int nv = 0, ns = 0, nh = 0, col = d;
  auto new_node = [&](int op, double arg, int a, int b, const Val& v) {
    Node n;
    n.op = op, n.arg = arg, n.a = a, n.b = b, n.inv = v.inv, n.leaf = v.leaf, n.name = v.name;
    nodes.push_back(n);
    return (int)nodes.size() - 1;
  };
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
        col_node.push_back(a.node);
        break;
    }
    if (op != EBN_OP_STORE) {
      Val& top = st.back();
      if (op == EBN_OP_INPUT && (int)arg >= d) top.node = col_node[(int)arg - d];   // a stored column IS its value
      else top.node = new_node(op, arg, kShape[op].pops >= 1 && kShape[op].pops <= 2 ? a.node : -1,
                               kShape[op].pops == 2 ? b.node : -1, top);
    }
  }
  const std::string result = in_eval(st.back());
  const std::string fails_body = expr_fails_codegen(nodes, st.back().node, d, setup, nh, eval, result);
  std::string s;
  s += "namespace ebn {\ntemplate <class Ops>\nstruct JitExpr {\n";
  s += "  static constexpr int ID = EBN_LS_EXPRESSION;\n";
  s += "  static constexpr bool HAS_FAILS = !Ops::strict;   // sign-only predicate for the failure count in fast mode\n";
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
  s += "    return " + result + ";\n  }\n";
  // g < 0 without the divisions and square roots whose VALUE does not matter for the sign (fast mode only; the
  // values eval() computes that the predicate does not read are dead code here and the compiler drops them)
  s += "  __device__ __forceinline__ static bool fails(const Ctx& c, double* x) {\n";
  s += eval;
  s += fails_body;
  s += "  }\n};\n}  // namespace ebn\n";
  return s;
}

}  // namespace ebn
