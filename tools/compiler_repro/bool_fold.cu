// Minimal reproducer of the boolean folding that made the first version of the generated sign-only predicate
// (csrc/ebn_expr.cu: FailsGen) wrong on the device and right on the host.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -ptx bool_fold.cu -o bool_fold.ptx    (CUDA 12.9; -O0 alike)
//   grep -c setp bool_fold.ptx        -> 2
//
// The kernel must compare all of a, b, c, d (five comparisons with the NaN test of k0); the PTX holds
//   setp.num.f64 p1, k0, k0;  setp.lt.f64 p2, a, 0;  and.pred p3, p1, p2
// i.e. out = !isnan(k0) && a < 0: the terms (b < 1.2), (c < 0), (d < 2.37) are gone. With k0 = -0.7, a = 0.7,
// b = 0.8 the right answer is 1 (g++ prints 1 for the same statements), the device answers 0.
// Any of these makes the folding correct again: isnan(k0) instead of (k0 != k0); && / || instead of & / |;
// a single NaN flag instead of two; negations in their own temporaries. The generator now keeps every truth
// value in an int behind an empty asm, which takes the question away from the optimiser.
__global__ void t1(const double* k, const double* X, int* out) {
  const double k0 = k[0];
  const double a = X[0], b = X[1], c = X[2], d = X[3];
  const bool p0 = (k0 != k0);
  const bool p1 = p0;
  const bool p2 = (k0 != k0);
  const bool p3 = p2;
  const bool p4 = p3;
  const bool p5 = p4;
  const bool p6 = (a < 0.0);
  const bool p7 = (b < 1.2);
  const bool p11 = (c < 0.0);
  const bool p12 = (d < 2.37);
  const bool p13 = ((p12 | p11) & !p3);
  const bool p14 = ((p13 | p7) & !p4);
  const bool p15 = ((p14 | p6) & !(p5 | p1));
  out[0] = p15;
}
