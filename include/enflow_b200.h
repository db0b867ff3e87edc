/* enflow_b200 C ABI: the drop-in boundary of the B200 hot path.
 *
 * The reference (bharath-raghavan/enflow) is pure Python/PyTorch and has no FFI; these entry points
 * are what a binding for its hot path would call.  Each one names the reference code it replaces
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it says "host"; floats are fp32, indices int32;
 *   - the caller owns all memory (inputs, outputs, workspace); the library never allocates or frees
 *     device memory and keeps no reference after a call returns (calls are stream-ordered);
 *   - `stream` is a cudaStream_t passed as void*; all work of a call is ordered after what was enqueued on it before
 *     the call and before what is enqueued after it.  enflow_flow_backward additionally runs its weight-gradient
 *     reductions on one library-owned non-blocking stream (created on first use, on the device current then: one
 *     process per GPU, one host thread per process, as in the reference), forked from and joined back into `stream`
 *     inside the call with events, so the call is capturable in a CUDA graph like any other;
 *   - return value 0 = ok; non-zero = failure, message from enflow_last_error() (thread-local);
 *   - `status` (int[1], device) receives OR-ed flags: 1 = edge capacity exceeded (E > E_cap),
 *     2 = the reference's id_mapping lookup would have raised IndexError (data/base.py:137),
 *     4 = a tensor-core kernel found its shared-memory window misaligned (nothing was computed),
 *     8 = dims.fc was set but some molecule is not provably in the fully connected regime (repeat with fc = 0);
 *   - molecules are contiguous: molecule m owns atoms [mol_off[m], mol_off[m+1]).
 *   - hidden width is fixed at 128 (example/train.yaml:19), node features nf <= 8.
 */
#ifndef ENFLOW_B200_H
#define ENFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t B;             /* molecules in the batch                                   */
    int32_t N;             /* atoms in the batch                                       */
    int32_t nf;            /* node features                                            */
    int32_t L;             /* coupling steps = number of EGCLs (dynamics n_iter)       */
    int32_t E_cap;         /* edge capacity per layer                                  */
    int32_t max_n;         /* atoms in the largest molecule                            */
    float dt;              /* leap-frog step in LJ time units                          */
    float coords_weight;   /* EGCL coords_weight (egcl.py:11), 1.0 in Main             */
    int32_t mode;          /* edge-MLP arithmetic: 0 = fp32 FFMA, 1 = tcgen05 bf16x3 split (fp32-accurate), 2 = tcgen05 bf16 */
    int32_t fc;            /* 1: the caller expects every molecule to be in the fully connected regime (Data.edges = all ordered
                              pairs, base.py:122-144 with box >> extent, r_cut >= extent): the list is built once by index
                              arithmetic and every coupling step only proves the regime on its positions; E_cap must be
                              sum n(n-1).  A molecule outside the regime sets status bit 8: repeat the pass with fc = 0. */
} enflow_dims_t;

const char* enflow_last_error(void);
int enflow_version(void);
int enflow_hidden(void);

/* ---- instrumentation (used by bench.py) ---------------------------------------------------------
 * enflow_launch_count: kernels launched by this library since the last reset.
 * enflow_timing_*: when enabled, the flow entry points bracket each kernel family with CUDA events on
 * the launch stream; enflow_timing_read sums the elapsed milliseconds per family (host arrays of
 * enflow_timing_kinds() entries; one family per kernel, order: edges (K0 or the fully connected list + its per-layer
 * check), node_pre, edge_fwd, run_sum (k_run_sum128), seg_cols (the column-grouped k_segment_sum128 behind dS),
 * seg_rows (row-grouped k_segment_sum128, FFMA mode), segment_sum3, node_post, coupling_fwd, coupling_bwd, coupling_inv,
 * edge_geom, edge_bwd (the backward edge kernel alone), edge_reduce, node_post_bwd, node_pre_bwd, col_perm, argmax, nll). */
long long enflow_launch_count(int reset);
int enflow_timing_enable(int enable);
int enflow_timing_kinds(void);
int enflow_timing_read(float* ms, int* counts);

/* ---- flat parameter buffer -------------------------------------------------------------------
 * All parameters of LFIntegrator(networks=L x EGCL, dequantize=ArgMax) live in ONE fp32 buffer
 * (gradients in a second one with the same layout, so data-parallel training needs one all-reduce).
 * enflow_param_layout writes, for the 15 EGCL tensors of every layer followed by the 4 ArgMax
 * tensors, the offset (in floats) and element count, in state_dict order:
 * networks.{i}.{edge_nn.0.weight, edge_nn.0.bias, edge_nn.2.weight, edge_nn.2.bias, node_nn.0.weight,
 * node_nn.0.bias, node_nn.2.weight, node_nn.2.bias, coord_nn.0.weight, coord_nn.0.bias,
 * coord_nn.2.weight, vel_scaling_nn.0.weight, .0.bias, .2.weight, .2.bias}, dequantize.network.{0,2}.{weight,bias}
 * (enflow/nn/egcl.py:21-55, enflow/nn/argmax.py:9-12, enflow/flow/base.py:8-9).
 * offsets/counts: host arrays of 15*L + 4 entries.  Returns the total buffer length in floats. */
int64_t enflow_param_layout(int nf, int L, int64_t* offsets, int64_t* counts);

/* ---- optimizer step on the flat buffers (torch.optim.Adam in enflow/main.py:177,222): in-place Adam over n floats,
 * step counter on the device (int[1], incremented by the call) so the launch is CUDA-graph capturable.  lr_dev
 * (nullable device float[1]) overrides lr: a captured graph then follows a per-batch StepLR (main.py:188,223). */
int enflow_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int32_t* step,
                     float lr, const float* lr_dev, float beta1, float beta2, float eps, void* stream);

/* ---- K0 neighbour list: Data.edges (enflow/data/base.py:122-144, utils/helpers.py:15-29) -------
 * pos/box are fp32 (pos_is_f64 = 0) or fp64 (1).  Output is grouped by row (CSR): row/col [E_cap],
 * rowptr [N+1]; ref_pos[e] (nullable) is the index the edge has in the reference's own ordering.
 * E_dev: int[2] = {min(E, E_cap), E}.  ws: int workspace of enflow_edges_workspace_ints(N). */
int64_t enflow_edges_workspace_ints(int N);
int enflow_build_edges(const void* pos, const void* box, int pos_is_f64, const float* r_cut, const int32_t* mol_off,
                       int B, int N, int E_cap, int32_t* row, int32_t* col, int32_t* rowptr, int32_t* ref_pos,
                       int32_t* E_dev, int32_t* status, int32_t* ws, void* stream);
/* column-grouped view of the same edges (needed by the backward scatter onto `col`): colptr [N+1], perm [E_cap] */
/* ---- fully connected regime (same reference semantics, base.py:122-144): when box >> extent and r_cut >= extent the
 * list is all ordered pairs per molecule, row-major.  enflow_fc_check proves that on the given fp32 positions (status
 * bit 8 if some molecule is not provably in the regime: conditions in csrc/fc.cu); enflow_fc_build writes the list by
 * index arithmetic: row/col [E_cap], rowptr [N+1], E_dev[2] as K0, optional column-grouped view colptr [N+1] / perm
 * [E_cap] as enflow_build_col_perm; eoff: B + 2 ints of scratch. */
int enflow_fc_check(const float* pos, const float* box, const float* r_cut, const int32_t* mol_off, int B, int32_t* status,
                    void* stream);
int enflow_fc_build(const int32_t* mol_off, int B, int N, int E_cap, int32_t* row, int32_t* col, int32_t* rowptr,
                    int32_t* E_dev, int32_t* colptr, int32_t* perm, int32_t* eoff, int32_t* status, void* stream);
int enflow_build_col_perm(const int32_t* col, const int32_t* rowptr, const int32_t* mol_off, int B, int N, int E_cap,
                          const int32_t* E_dev, int32_t* colptr, int32_t* perm, int32_t* ws, void* stream);

/* ---- K2 segmented reductions: unsorted_segment_sum / _mean (enflow/utils/helpers.py:54-70) ------
 * x [E,128] (or [E,3]) in CSR order, ptr [N+1]; perm (nullable) gathers rows x[perm[e]] instead.
 * apply_silu: reduce silu(x).  mean: divide by max(deg,1).  accumulate: out += result. */
int enflow_segment_sum128(const float* x, const int32_t* ptr, const int32_t* perm, int N, int E_cap, int apply_silu,
                          float* out, void* stream);
int enflow_segment_sum3(const float* x, const int32_t* ptr, const int32_t* perm, int N, int E_cap, int mean,
                        float scale, int accumulate, float* out, void* stream);

/* ---- EGCL pieces (enflow/nn/egcl.py:57-93); layer_params points at one layer inside the flat buffer */
int64_t enflow_pack_floats(int nf);
int enflow_pack_layer(const float* layer_params, int nf, float* packed, void* stream);
int enflow_node_pre_fwd(const float* h, int N, int nf, const float* layer_params, float* P, float* S, float* Q,
                        void* stream);
int enflow_edge_fwd(const int32_t* row, const int32_t* col, const int32_t* E_dev, int E_cap, const float* pos,
                    const float* box, const float* P, const float* S, const float* layer_params, const float* packed,
                    int nf, float* wr_scratch, float* z2, float* z3, float* s, float* trans, void* stream);
/* tensor-core variant of enflow_edge_fwd (tcgen05.mma, TMEM accumulators). wimg: enflow_tc_pack_bytes() bytes
 * written by enflow_tc_pack_layer (swizzled bf16 hi/lo images of edge_nn.2 / coord_nn.0 weights), 16-byte aligned.
 * mode 1: bf16x3 operand split, fp32-accurate; mode 2: plain bf16 operands.
 * Instead of storing edge_attr [E,128] it reduces it over each row inside the kernel into per-"run" partials
 * (a run = the edges of one row inside one 16-edge block): runs [enflow_run_rows(E_cap,N)][128], indexed through
 * mis [N+2] from enflow_run_index; enflow_run_sum128 adds the runs of each row in order -> unsorted_segment_sum
 * (enflow/utils/helpers.py:54-60) of edge_attr, deterministic. */
int64_t enflow_tc_pack_bytes(void);
int enflow_tc_pack_layer(const float* layer_params, int nf, void* wimg, void* stream);
int64_t enflow_run_rows(int E_cap, int N);
int64_t enflow_run_scratch_ints(int N);
int enflow_run_index(const int32_t* rowptr, int N, int32_t* mis, int32_t* scratch, void* stream);
int enflow_run_sum128(const float* runs, const int32_t* rowptr, const int32_t* mis, int N, int E_cap, float* out,
                      void* stream);
int enflow_edge_fwd_tc(int mode, const int32_t* row, const int32_t* col, const int32_t* E_dev, int E_cap,
                       const float* pos, const float* box, const float* P, const float* S, const float* layer_params,
                       const void* wimg, int nf, const int32_t* rowptr, const int32_t* mis, float* runs, float* s,
                       float* trans, void* stream);
int enflow_node_post_fwd(const float* h, const float* agg, int N, int nf, const float* layer_params,
                         const float* packed, float* z4, float* G, void* stream);

/* ---- K3 coupling: LFIntegrator.forward/reverse body (enflow/flow/dynamics.py:14-21, 27-33) ---- */
int enflow_coupling_fwd(const float* Q, const float* F, const float* G, const float* h, const float* g,
                        const float* pos, const float* vel, const float* box, const int32_t* mol_off, int B, int nf,
                        float dt, float* h_out, float* g_out, float* pos_out, float* vel_out, float* ldj_mol,
                        void* stream);
int enflow_coupling_bwd(const float* Q, const float* vel_in, const float* dldj, int N, int nf, float dt, float* dh,
                        float* dg, float* dpos, float* dvel, float* dQ, float* dF, float* dG, void* stream);
int enflow_coupling_inv_pre(const float* g, const float* vel, const float* box, int N, int nf, float dt, float* h,
                            float* pos, void* stream);
int enflow_coupling_inv_post(const float* Q, const float* F, const float* G, const int32_t* mol_off, int B, int nf,
                             float dt, float* g, float* vel, float* neg_ldj_mol, void* stream);

/* ---- K4 dequantiser: ArgMax.forward (enflow/nn/argmax.py:14-26); eps is the injected N(0,1) noise */
int enflow_argmax_fwd(const float* h, const float* eps, int N, int nf, const float* argmax_params,
                      const int32_t* mol_off, int B, float* z, float* logq_atom, double* logq_mol, float* log_q,
                      void* stream);

/* ---- K5 likelihood: Alchemical_NLL (enflow/flow/loss.py:11-25) --------------------------------
 * A molecule is evaluated in slices of 128 atoms (one CTA each): `mol_term` is caller-provided scratch of
 * B * enflow_nll_slices(max_n) doubles (one partial per molecule and slice, added in index order). */
int enflow_nll_slices(int max_n);
int enflow_nll_fwd(const float* pos, const float* vel, const float* h, const float* g, const int32_t* mol_off, int B,
                   int N, int nf, int max_n, float kBT, float softening, float z_lj, const float* ldj,
                   double* mol_term, float* loss, void* stream);
int enflow_nll_bwd(const float* pos, const float* vel, const float* h, const float* g, const int32_t* mol_off, int B,
                   int nf, int max_n, float kBT, float softening, const float* dloss, float* dpos, float* dvel,
                   float* dh, float* dg, float* dldj, void* stream);

/* ---- prior sampler for generate: the soft Lennard-Jones Langevin run that the reference does through
 * OpenMM (enflow/data/lj.py:32-89 potential and cutoff, enflow/data/simulated.py:109-132 LangevinMiddleIntegrator,
 * minimisation, Maxwell velocities), in the reduced units of the likelihood (sigma = eps = mass = 1).
 * pos, vel, force: device fp64 [N,3]; box: HOST double[3] (periodic cell); ws: enflow_lj_prior_workspace_doubles(N)
 * device doubles; energy: device double[2] = {potential, kinetic} or NULL.
 * run: n_steps of  v += dt f; x += dt/2 v; v = a v + sqrt(kBT (1 - a^2)) N(0,1); x += dt/2 v  with a = exp(-friction dt);
 * noise = Philox4x32-10(seed; step0 + s, atom): the same (seed, step0) reproduces the trajectory bit for bit. */
int64_t enflow_lj_prior_workspace_doubles(int N);
int enflow_lj_prior_forces(const double* pos, int N, const double* box, double softening, double cutoff, double* ws,
                           double* force, double* energy, void* stream);
int enflow_lj_prior_minimize(double* pos, int N, const double* box, double softening, double cutoff, int iters,
                             double rate, double cap, double* ws, void* stream);
int enflow_lj_prior_velocities(double* vel, int N, double kBT, uint64_t seed, void* stream);
int enflow_lj_prior_run(double* pos, double* vel, int N, const double* box, double softening, double cutoff, double dt,
                        double a, double kBT, int n_steps, uint64_t seed, uint64_t step0, double* ws, double* energy,
                        void* stream);

/* ---- whole flow: LFIntegrator.forward / its backward / .reverse (enflow/flow/dynamics.py:10-37) --
 * One call enqueues every kernel of the pass on `stream`; no host synchronisation inside.
 * workspace: enflow_flow_workspace_bytes(dims, training) bytes, reused by the matching backward.
 * eps: [N,nf] ArgMax noise, or NULL to skip dequantisation (h_in is then used as is).
 * Outputs: state after L steps, ldj_mol [B] = per-molecule sum of Q, ldj [1] = log_q + sum ldj_mol. */
size_t enflow_flow_workspace_bytes(const enflow_dims_t* dims, int training);
int enflow_flow_forward(const enflow_dims_t* dims, const float* params, const float* h_in, const float* g_in,
                        const float* pos_in, const float* vel_in, const float* box, const float* r_cut,
                        const int32_t* mol_off, const float* eps, void* workspace, size_t workspace_bytes,
                        int training, float* h_out, float* g_out, float* pos_out, float* vel_out, float* ldj_mol,
                        float* ldj, int32_t* status, void* stream);
/* d*_out: gradients w.r.t. the forward outputs (consumed/overwritten); dldj [1]; grads: flat buffer, accumulated (+=) */
int enflow_flow_backward(const enflow_dims_t* dims, const float* params, float* grads, const float* h_in,
                         const float* box, const int32_t* mol_off, const float* eps, void* workspace,
                         size_t workspace_bytes, float* dh_out, float* dg_out, float* dpos_out, float* dvel_out,
                         const float* dldj, int32_t* status, void* stream);
/* in-place inverse; neg_ldj_mol [B] (nullable) receives -sum(Q) per molecule; quantize: apply ArgMax.reverse */
int enflow_flow_reverse(const enflow_dims_t* dims, const float* params, float* h, float* g, float* pos, float* vel,
                        const float* box, const float* r_cut, const int32_t* mol_off, void* workspace,
                        size_t workspace_bytes, int quantize, float* neg_ldj_mol, int32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif
