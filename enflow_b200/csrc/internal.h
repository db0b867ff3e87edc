// Internal launch functions (one per kernel family); the extern "C" surface is in api.cu.
#pragma once
#include "common.cuh"

int64_t enf_edges_workspace_ints(int N);
template <typename T>
int enf_build_edges_t(const T* pos, const T* box, const float* r_cut, const int* mol_off, int B, int N, int E_cap,
                      int* row, int* col, int* rowptr, int* ref_pos, int* E_dev, int* status, int* ws,
                      cudaStream_t st);
int enf_build_col_perm(const int* col, const int* rowptr, const int* mol_off, int B, int N, int E_cap,
                       const int* E_dev, int* colptr, int* perm, int* ws, const int* skip, cudaStream_t st);
int enf_edges_same(const int* row_a, const int* col_a, const int* E_a, const int* row_b, const int* col_b,
                   const int* E_b, int* same, cudaStream_t st);

// implicit all-pairs lists in the fully connected regime (fc.cu)
int enf_fc_check(const float* pos, const float* box, const float* r_cut, const int* mol_off, int B, int* status,
                 cudaStream_t st);
int enf_fc_build(const int* mol_off, int B, int N, int E_cap, int* row, int* col, int* rowptr, int* E_dev, int* colptr,
                 int* perm, int* eoff, int* status, cudaStream_t st);

int enf_segment_sum128(const float* x, const int* ptr, const int* perm, int N, int E_cap, int apply_silu, float* out,
                       cudaStream_t st);
int enf_segment_sum3(const float* x, const int* ptr, const int* perm, int N, int E_cap, int mean, float scale,
                     int accumulate, float* out, cudaStream_t st);

int enf_pack_layer(const float* layer_params, int nf, float* packed, cudaStream_t st);
int enf_node_pre_fwd(const float* h, int N, int nf, const float* lp, float* P, float* S, float* Q, cudaStream_t st);
int64_t enf_node_pre_partial_floats(int N, int nf);
int enf_node_pre_bwd(const float* h, int N, int nf, const float* lp, const float* dP, const float* dS,
                     const float* dQ, float* dh, float* lgrad, float* partial, cudaStream_t st, cudaStream_t st_red);
int enf_node_post_fwd(const float* h, const float* agg, int N, int nf, const float* lp, const float* packed,
                      float* z4, float* G, cudaStream_t st);
int64_t enf_node_post_partial_floats(int N, int nf);
int enf_node_post_bwd(const float* h, const float* agg, const float* z4, const float* dG, int N, int nf,
                      const float* lp, const float* packed, float* dagg, float* dh, float* lgrad, float* partial,
                      cudaStream_t st);

int64_t enf_edge_partial_floats();
int enf_edge_fwd(const int* row, const int* col, const int* E_dev, int E_cap, const float* pos, const float* box,
                 const float* P, const float* S, const float* lp, const float* packed, int nf, float* wr, float* z2,
                 float* z3, float* s_out, float* trans, cudaStream_t st);
int enf_edge_bwd(const int* row, const int* col, const int* rowptr, const int* E_dev, int E_cap, const float* pos,
                 const float* box, const float* P, const float* S, const float* lp, int nf, float* wr,
                 const float* z2, const float* z3, const float* s_saved, const float* dagg, const float* dF,
                 float coords_weight, float* dz1, float* dd, float* lgrad, float* partial, cudaStream_t st);

int enf_coupling_fwd(const float* Q, const float* F, const float* G, const float* h, const float* g,
                     const float* pos, const float* vel, const float* box, const int* mol_off, int B, int nf,
                     float dt, float* h_o, float* g_o, float* pos_o, float* vel_o, float* ldj_mol, cudaStream_t st);
int enf_coupling_bwd(const float* Q, const float* vel_in, const float* dldj, int N, int nf, float dt, float* dh,
                     float* dg, float* dpos, float* dvel, float* dQ, float* dF, float* dG, cudaStream_t st);
int enf_coupling_inv_pre(const float* g, const float* vel, const float* box, int N, int nf, float dt, float* h,
                         float* pos, cudaStream_t st);
int enf_coupling_inv_post(const float* Q, const float* F, const float* G, const int* mol_off, int B, int nf,
                          float dt, float* g, float* vel, float* neg_ldj_mol, cudaStream_t st);

int enf_argmax_fwd(const float* h, const float* eps, int N, int nf, const float* ap, const int* mol_off, int B,
                   float* z, float* logq_atom, double* logq_mol, float* log_q, cudaStream_t st);
int64_t enf_argmax_partial_floats(int N, int nf);
int enf_argmax_bwd(const float* h, const float* eps, int N, int nf, const float* ap, const float* dz,
                   const float* dlogq, float* agrad, float* partial, cudaStream_t st);
int enf_argmax_reverse(float* h, int N, int nf, cudaStream_t st);

int enf_nll_slices(int max_n);
int enf_nll_fwd(const float* pos, const float* vel, const float* h, const float* g, const int* mol_off, int B, int N,
                int nf, int max_n, float kBT, float softening, float z_lj, const float* ldj, double* mol_term,
                float* loss, cudaStream_t st);
int enf_nll_bwd(const float* pos, const float* vel, const float* h, const float* g, const int* mol_off, int B, int nf,
                int max_n, float kBT, float softening, const float* dloss, float* dpos, float* dvel, float* dh,
                float* dg, float* dldj, cudaStream_t st);
int enf_ldj_total(const float* ldj_mol, int B, const float* log_q, float* ldj, cudaStream_t st);

// prior sampler for generate (lj_prior.cu): periodic soft-LJ forces + LangevinMiddle steps, fp64, host box[3]
int64_t enf_lj_prior_workspace_doubles(int N);
int enf_lj_prior_forces(const double* pos, int N, const double* box, double soft, double rc, double* ws, double* force,
                        double* energy, cudaStream_t st);
int enf_lj_prior_minimize(double* pos, int N, const double* box, double soft, double rc, int iters, double rate,
                          double cap, double* ws, cudaStream_t st);
int enf_lj_prior_velocities(double* vel, int N, double kBT, uint64_t seed, cudaStream_t st);
int enf_lj_prior_run(double* pos, double* vel, int N, const double* box, double soft, double rc, double dt, double a,
                     double kBT, int n_steps, uint64_t seed, uint64_t step0, double* ws, double* energy,
                     cudaStream_t st);

// tensor-core (tcgen05) edge kernels; mode 1 = bf16x3 split (fp32-accurate), mode 2 = bf16
int64_t enf_tc_pack_bytes();
int enf_tc_pack_layer(const float* lp, int nf, unsigned char* img, cudaStream_t st);
int enf_tc_pack_layers(const float* lp0, int nf, int L, int64_t param_stride, unsigned char* img0, int64_t img_stride,
                       cudaStream_t st);
int enf_edge_fwd_tc(int mode, const int* row, const int* col, const int* E_dev, int E_cap, const float* pos,
                    const float* box, const float* P, const float* S, const float* lp, const unsigned char* wimg,
                    int nf, const int* rowptr, const int* mis, float* runs, float* s_out, float* trans,
                    cudaStream_t st);
int enf_edge_bwd_tc_geom(const int* row, const int* col, const int* rowptr, const int* E_dev, int E_cap, const float* pos,
                         const float* box, const float* s_saved, const float* dF, float coords_weight, const int* mis,
                         unsigned char* geom, cudaStream_t st);
int enf_edge_bwd_tc(int mode, const int* E_dev, int E_cap, const float* P, const float* S, const float* lp,
                    const unsigned char* wimg, int nf, const float* dagg, float* runs, float* dz1, float* dd, float* lgrad,
                    float* partial, unsigned char* geom, int* status, cudaStream_t st, cudaStream_t st_red);
int64_t enf_edge_bwd_geom_bytes(int E_cap);
// tensor-core node_model (node_tc.cu); weight images live behind the edge images in the same per-layer buffer
int enf_node_post_fwd_tc(int mode, const float* h, const float* agg, int N, int nf, const float* lp,
                         const unsigned char* wimg, float* z4, float* G, cudaStream_t st);
int enf_node_post_bwd_tc(int mode, const float* h, const float* agg, const float* z4, const float* dG, int N, int nf,
                         const float* lp, const unsigned char* wimg, float* dagg, float* dh, float* lgrad,
                         float* partial, cudaStream_t st, cudaStream_t st_red);
// per-run partial segment sums (segment.cu): mis = N+2 ints, runs = enf_run_rows(E_cap, N) x 128 floats
int64_t enf_scan_scratch_ints(int64_t n);
int enf_run_index(const int* rowptr, int N, int* mis, int* scratch, cudaStream_t st);
int enf_run_sum128(const float* runs, const int* rowptr, const int* mis, int N, int E_cap, float* out, cudaStream_t st);
int enf_run_sum128_sum3(const float* runs, const int* rowptr, const int* mis, int N, int E_cap, float* out,
                        const float* x3, int mean3, float scale3, int accumulate3, float* out3, cudaStream_t st);
int enf_segment_sum128_sum3(const float* x, const int* ptr, const int* perm, int N, int E_cap, int apply_silu, float* out,
                            const float* x3, float scale3, float* out3, const int* rowptr3, cudaStream_t st);
inline int64_t enf_run_rows(int E_cap, int N) { return (int64_t)E_cap / 16 + N + 2; }
int enf_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int* step, float lr, const float* lr_dev,
                  float beta1, float beta2, float eps, cudaStream_t st);
