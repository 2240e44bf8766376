// Fused constraint evaluation on the device.
//
// Replaces halo2_proofs v2023_04_20 (un-vendored; /root/reference/Cargo.toml:21-25)
// `plonk::evaluation::Evaluator::evaluate_h` (GraphEvaluator for the custom gates, then the
// permutation and lookup terms, all folded by y in that order) and the `compress_expressions` step of
// `plonk::lookup::prover::commit_permuted`; call site /root/reference/src/wnn.rs:242-259.
//
// Gate / lookup polynomials arrive as RPN programs (host front-end serialises the compressed-selector
// constraint system); a thread owns one row of the (extended) domain, evaluates every program with a
// small register/local stack and folds the results with Horner in y.  One thread per row keeps every
// column read coalesced across the warp (rotations are a constant row offset).
#include "expr.cuh"

namespace zg {

namespace {

constexpr int EX_THREADS = 128;

__device__ __forceinline__ Fr ldf(const Fr* p) {
  Fr r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void stf(Fr* p, const Fr& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// row `idx` rotated by `rot` rows of the 2^k domain: a shift by rot * scale inside its block of mask + 1 rows
__device__ __forceinline__ uint32_t rot_idx(uint32_t idx, int32_t rot, uint32_t scale, uint32_t mask) {
  return (idx & ~mask) | ((uint32_t)((int32_t)idx + rot * (int32_t)scale) & mask);
}

__device__ Fr eval_program(const ExprEnv& E, uint32_t begin, uint32_t end, uint32_t idx) {
  Fr st[EXPR_STACK];
  int sp = 0;
  for (uint32_t pc = begin; pc < end; pc++) {
    const uint32_t w = E.ops[pc];
    const uint32_t op = w & 0xff, arg = w >> 8;
    switch (op) {
      case OP_CONST:
        st[sp++] = ldf(E.constants + arg);
        break;
      case OP_ADVICE:
      case OP_FIXED:
      case OP_INSTANCE: {
        const int kind = (int)op - (int)OP_ADVICE;
        const Fr* col = E.cols[kind][E.qcol[kind][arg]];
        st[sp++] = ldf(col + rot_idx(idx, E.qrot[kind][arg], E.rot_scale, E.wrap_mask));
        break;
      }
      case OP_NEG:
        st[sp - 1] = fp_neg(st[sp - 1]);
        break;
      case OP_ADD:
        st[sp - 2] = fp_add(st[sp - 2], st[sp - 1]);
        sp--;
        break;
      case OP_SUB:
        st[sp - 2] = fp_sub(st[sp - 2], st[sp - 1]);
        sp--;
        break;
      case OP_MUL:
        st[sp - 2] = fp_mul(st[sp - 2], st[sp - 1]);
        sp--;
        break;
      case OP_SCALE:
        st[sp - 1] = fp_mul(st[sp - 1], ldf(E.constants + arg));
        break;
      default:
        break;
    }
  }
  return st[0];
}

// theta-Horner over programs [first, first+count): acc = acc * theta + expr
__device__ Fr compress(const ExprEnv& E, const uint32_t* prog_off, uint32_t first, uint32_t count, const Fr& theta, uint32_t idx) {
  Fr acc = eval_program(E, prog_off[first], prog_off[first + 1], idx);
  for (uint32_t p = first + 1; p < first + count; p++)
    acc = fp_add(fp_mul(acc, theta), eval_program(E, prog_off[p], prog_off[p + 1], idx));
  return acc;
}

__global__ void __launch_bounds__(EX_THREADS) k_compress_lookups(ExprEnv E, LookupProgs lp, Fr theta, Fr* out_in, Fr* out_tab,
                                                                 size_t out_stride) {
  uint32_t idx = E.row0 + blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t l = blockIdx.y;
  if (idx >= E.size) return;
  stf(out_in + l * out_stride + idx, compress(E, lp.prog_off, lp.in_first[l], lp.in_count[l], theta, idx));
  stf(out_tab + l * out_stride + idx, compress(E, lp.prog_off, lp.tab_first[l], lp.tab_count[l], theta, idx));
}

__global__ void __launch_bounds__(EX_THREADS) k_h_gates(ExprEnv E, const uint32_t* prog_off, uint32_t nprogs, Fr y, Fr* h) {
  uint32_t idx = E.row0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= E.size) return;
  Fr acc = fp_zero<FrParams>();
  for (uint32_t p = 0; p < nprogs; p++) acc = fp_add(fp_mul(acc, y), eval_program(E, prog_off[p], prog_off[p + 1], idx));
  stf(h + idx, acc);
}

__global__ void __launch_bounds__(EX_THREADS) k_h_permutation(PermEnv P, Fr beta, Fr gamma, Fr y, Fr delta, Fr* h) {
  uint32_t idx = P.row0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P.size) return;
  const uint32_t r_next = rot_idx(idx, 1, P.rot_scale, P.wrap_mask);
  const uint32_t r_last = rot_idx(idx, P.last_rot, P.rot_scale, P.wrap_mask);
  const Fr one = fp_one<FrParams>();
  Fr acc = ldf(h + idx);
  const Fr l0 = ldf(P.l0 + idx), l_last = ldf(P.l_last + idx), l_active = ldf(P.l_active + idx);
  // l_0 * (1 - z_0)
  Fr z_first = ldf(P.z_cosets[0] + idx);
  acc = fp_add(fp_mul(acc, y), fp_mul(fp_sub(one, z_first), l0));
  // l_last * (z_l^2 - z_l)
  Fr z_lastset = ldf(P.z_cosets[P.nsets - 1] + idx);
  acc = fp_add(fp_mul(acc, y), fp_mul(fp_sub(fp_sqr(z_lastset), z_lastset), l_last));
  // l_0 * (z_i - z_{i-1}(w^last X))
  for (uint32_t s = 1; s < P.nsets; s++) {
    Fr zi = ldf(P.z_cosets[s] + idx);
    Fr zp = ldf(P.z_cosets[s - 1] + r_last);
    acc = fp_add(fp_mul(acc, y), fp_mul(fp_sub(zi, zp), l0));
  }
  // (1 - (l_last + l_blind)) * (z_i(wX) prod (p + beta sigma + gamma) - z_i(X) prod (p + delta^j beta X + gamma))
  Fr cur_delta = fp_mul(beta, ldf(P.coset_x + idx));
  for (uint32_t s = 0; s < P.nsets; s++) {
    Fr left = ldf(P.z_cosets[s] + r_next);
    Fr right = ldf(P.z_cosets[s] + idx);
    uint32_t c0 = s * P.chunk, c1 = c0 + P.chunk < P.m ? c0 + P.chunk : P.m;
    for (uint32_t c = c0; c < c1; c++) {
      Fr v = ldf(P.col_cosets[c] + idx);
      Fr sg = ldf(P.sigma_cosets[c] + idx);
      left = fp_mul(left, fp_add(fp_add(v, fp_mul(beta, sg)), gamma));
      right = fp_mul(right, fp_add(fp_add(v, cur_delta), gamma));
      cur_delta = fp_mul(cur_delta, delta);
    }
    acc = fp_add(fp_mul(acc, y), fp_mul(fp_sub(left, right), l_active));
  }
  stf(h + idx, acc);
}

__global__ void __launch_bounds__(EX_THREADS) k_h_lookup(ExprEnv E, LookupProgs lp, uint32_t l, LookupHEnv L, Fr theta, Fr beta,
                                                         Fr gamma, Fr y, Fr* h) {
  uint32_t idx = E.row0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= E.size) return;
  const uint32_t r_next = rot_idx(idx, 1, E.rot_scale, E.wrap_mask);
  const uint32_t r_prev = rot_idx(idx, -1, E.rot_scale, E.wrap_mask);
  const Fr one = fp_one<FrParams>();
  Fr ci = compress(E, lp.prog_off, lp.in_first[l], lp.in_count[l], theta, idx);
  Fr ct = compress(E, lp.prog_off, lp.tab_first[l], lp.tab_count[l], theta, idx);
  Fr table_value = fp_mul(fp_add(ci, beta), fp_add(ct, gamma));
  Fr z = ldf(L.z + idx), z_next = ldf(L.z + r_next);
  Fr a = ldf(L.a + idx), a_prev = ldf(L.a + r_prev), s = ldf(L.s + idx);
  Fr l0 = ldf(L.l0 + idx), l_last = ldf(L.l_last + idx), l_active = ldf(L.l_active + idx);
  Fr a_minus_s = fp_sub(a, s);
  Fr acc = ldf(h + idx);
  acc = fp_add(fp_mul(acc, y), fp_mul(fp_sub(one, z), l0));
  acc = fp_add(fp_mul(acc, y), fp_mul(fp_sub(fp_sqr(z), z), l_last));
  Fr t = fp_sub(fp_mul(fp_mul(z_next, fp_add(a, beta)), fp_add(s, gamma)), fp_mul(z, table_value));
  acc = fp_add(fp_mul(acc, y), fp_mul(t, l_active));
  acc = fp_add(fp_mul(acc, y), fp_mul(a_minus_s, l0));
  acc = fp_add(fp_mul(acc, y), fp_mul(fp_mul(a_minus_s, fp_sub(a, a_prev)), l_active));
  stf(h + idx, acc);
}

inline uint32_t nblk(uint32_t n) { return (n + EX_THREADS - 1) / EX_THREADS; }

}  // namespace

void expr_compress_lookups(const ExprEnv& env, const LookupProgs& lp, uint32_t n_lookups, const Fr& theta, Fr* out_in, Fr* out_tab,
                           size_t out_stride, cudaStream_t st, LaunchCounter lc) {
  if (!n_lookups) return;
  k_compress_lookups<<<dim3(nblk(env.size - env.row0), n_lookups), EX_THREADS, 0, st>>>(env, lp, theta, out_in, out_tab, out_stride);
  lc++;
}
void expr_h_gates(const ExprEnv& env, const uint32_t* prog_off, uint32_t n_gate_progs, const Fr& y, Fr* h, cudaStream_t st,
                  LaunchCounter lc) {
  k_h_gates<<<nblk(env.size - env.row0), EX_THREADS, 0, st>>>(env, prog_off, n_gate_progs, y, h);
  lc++;
}
void expr_h_permutation(const PermEnv& pe, const Fr& beta, const Fr& gamma, const Fr& y, const Fr& delta, Fr* h, cudaStream_t st,
                        LaunchCounter lc) {
  if (!pe.nsets) return;
  k_h_permutation<<<nblk(pe.size - pe.row0), EX_THREADS, 0, st>>>(pe, beta, gamma, y, delta, h);
  lc++;
}
void expr_h_lookup(const ExprEnv& env, const LookupProgs& lp, uint32_t l, const LookupHEnv& le, const Fr& theta, const Fr& beta,
                   const Fr& gamma, const Fr& y, Fr* h, cudaStream_t st, LaunchCounter lc) {
  k_h_lookup<<<nblk(env.size - env.row0), EX_THREADS, 0, st>>>(env, lp, l, le, theta, beta, gamma, y, h);
  lc++;
}

}  // namespace zg
