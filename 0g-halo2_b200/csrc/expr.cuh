// Constraint-expression evaluation on the device: lookup compression over the 2^k domain and the
// fused quotient numerator (Evaluator::evaluate_h) over the extended coset.
#pragma once
#include <cuda_runtime.h>
#include "poly.cuh"

namespace zg {

// RPN op word: opcode in bits 0..7, argument in bits 8..31
enum : uint32_t { OP_CONST = 0, OP_ADVICE = 1, OP_FIXED = 2, OP_INSTANCE = 3, OP_NEG = 4, OP_ADD = 5, OP_MUL = 6, OP_SCALE = 7, OP_SUB = 8 };
constexpr int EXPR_STACK = 12;

// Everything an expression needs to read a query (col, rotation) at a row.  All pointers are device
// pointers; `cols[kind]` is a device array of column base pointers of `size` elements each.
struct ExprEnv {
  const Fr* const* cols[3];      // [advice, fixed, instance] -> column pointers
  const uint32_t* qcol[3];       // query -> column index
  const int32_t* qrot[3];        // query -> rotation
  const Fr* constants;           // constant pool (Montgomery)
  const uint32_t* ops;           // all programs, concatenated
  uint32_t row0 = 0;             // the launch covers rows [row0, size): a rank's coset block when a proof is spread over GPUs
  uint32_t size;                 // end row: 2^k, or cosets * B on the extended domain (extdomain.cuh)
  uint32_t rot_scale;            // 1 on the 2^k domain, B / n on the extended domain
  uint32_t wrap_mask;            // rotations wrap inside blocks of wrap_mask + 1 rows (2^k - 1, or B - 1)
};

// lookup compression: for lookup l (grid.y), out_in[l][row], out_tab[l][row] = theta-Horner of its programs.
// prog ranges: lookup l's input programs are [in_begin[l], in_begin[l+1]) in the program table, same for tables.
struct LookupProgs {
  const uint32_t* prog_off;      // program p = ops[prog_off[p] .. prog_off[p+1])
  const uint32_t* in_first;      // per lookup: first input program, count
  const uint32_t* in_count;
  const uint32_t* tab_first;
  const uint32_t* tab_count;
};
void expr_compress_lookups(const ExprEnv& env, const LookupProgs& lp, uint32_t n_lookups, const Fr& theta, Fr* out_in,
                           Fr* out_tab, size_t out_stride, cudaStream_t st, LaunchCounter lc);

// quotient numerator, gates part: h[idx] = Horner_y over gate programs [0, n_gate_progs) (h overwritten)
void expr_h_gates(const ExprEnv& env, const uint32_t* prog_off, uint32_t n_gate_progs, const Fr& y, Fr* h, cudaStream_t st,
                  LaunchCounter lc);

struct PermEnv {
  const Fr* const* z_cosets;     // nsets
  const Fr* const* col_cosets;   // m permutation columns' cosets (advice/fixed/instance resolved by the host)
  const Fr* const* sigma_cosets; // m
  const Fr* l0;
  const Fr* l_last;
  const Fr* l_active;
  const Fr* coset_x;             // X = zeta * ext_omega^idx
  uint32_t nsets, m, chunk, size, rot_scale, wrap_mask;
  uint32_t row0 = 0;             // rows [row0, size)
  int32_t last_rot;              // -(blinding_factors + 1)
};
void expr_h_permutation(const PermEnv& pe, const Fr& beta, const Fr& gamma, const Fr& y, const Fr& delta, Fr* h,
                        cudaStream_t st, LaunchCounter lc);

struct LookupHEnv {
  const Fr* z;
  const Fr* a;
  const Fr* s;
  const Fr* l0;
  const Fr* l_last;
  const Fr* l_active;
};
// one lookup's five terms folded into h; input/table compressed on the fly from programs of lookup `l`
void expr_h_lookup(const ExprEnv& env, const LookupProgs& lp, uint32_t l, const LookupHEnv& le, const Fr& theta, const Fr& beta,
                   const Fr& gamma, const Fr& y, Fr* h, cudaStream_t st, LaunchCounter lc);

}  // namespace zg
