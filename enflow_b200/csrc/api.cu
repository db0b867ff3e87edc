// extern "C" surface of libenflow_b200.so (declared in include/enflow_b200.h).
#include <stdarg.h>
#include <string.h>

#include "internal.h"

static thread_local char g_err[512] = "";

void enf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#define ST(s) reinterpret_cast<cudaStream_t>(s)

// ---- instrumentation: launch counter and optional CUDA-event timing per kernel family ------------
#include <vector>
static long long g_launches = 0;
void enf_count_launch() { ++g_launches; }

struct TimingSlot { cudaEvent_t a, b; int kind; };
static std::vector<TimingSlot> g_slots;
static size_t g_used = 0;
static bool g_timing = false;

void enf_time_begin(int kind, cudaStream_t st) {
    if (!g_timing) return;
    if (g_used == g_slots.size()) {
        TimingSlot s;
        cudaEventCreate(&s.a);
        cudaEventCreate(&s.b);
        g_slots.push_back(s);
    }
    g_slots[g_used].kind = kind;
    cudaEventRecord(g_slots[g_used].a, st);
}
void enf_time_end(cudaStream_t st) {
    if (!g_timing) return;
    cudaEventRecord(g_slots[g_used].b, st);
    ++g_used;
}
bool enf_timing_on() { return g_timing; }

// ---- side streams and the events that order them against the caller's stream
static cudaStream_t g_side[2] = {nullptr, nullptr};
static cudaEvent_t g_chain_ev[64];
static cudaEvent_t g_mark_ev[8];
static int g_chain_next = 0;
static bool g_side_made = false;
static void side_make() {
    if (g_side_made) return;
    for (int i = 0; i < 2; ++i) cudaStreamCreateWithFlags(&g_side[i], cudaStreamNonBlocking);
    for (int i = 0; i < 64; ++i) cudaEventCreateWithFlags(&g_chain_ev[i], cudaEventDisableTiming);
    for (int i = 0; i < 8; ++i) cudaEventCreateWithFlags(&g_mark_ev[i], cudaEventDisableTiming);
    g_side_made = true;
}
cudaStream_t enf_side_stream(int which) {
    side_make();
    return g_side[which & 1];
}
void enf_mark(int id, cudaStream_t st) {
    side_make();
    cudaEventRecord(g_mark_ev[id & 7], st);
}
void enf_wait_mark(int id, cudaStream_t st) {
    side_make();
    cudaStreamWaitEvent(st, g_mark_ev[id & 7], 0);
}
void enf_chain(cudaStream_t from, cudaStream_t to) {
    if (from == to) return;
    side_make();
    cudaEvent_t ev = g_chain_ev[g_chain_next];       // a wait refers to the record made just before it: reuse is safe
    g_chain_next = (g_chain_next + 1) & 63;
    cudaEventRecord(ev, from);
    cudaStreamWaitEvent(to, ev, 0);
}

#pragma GCC visibility push(default)
extern "C" {

const char* enflow_last_error(void) { return g_err; }
int enflow_version(void) { return 100; }
int enflow_hidden(void) { return ENF_H; }

long long enflow_launch_count(int reset) {
    const long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}
// enable != 0: start collecting (drops earlier samples); enable == 0: stop
int enflow_timing_enable(int enable) {
    g_timing = enable != 0;
    if (enable) g_used = 0;
    return ENF_OK;
}
// host arrays of enflow_timing_kinds() entries: summed milliseconds and launch-group counts per family.
// Synchronises on the recorded events.
int enflow_timing_kinds(void) { return TK_COUNT; }
int enflow_timing_read(float* ms, int* counts) {
    for (int k = 0; k < TK_COUNT; ++k) { ms[k] = 0.f; counts[k] = 0; }
    for (size_t i = 0; i < g_used; ++i) {
        float t = 0.f;
        if (cudaEventSynchronize(g_slots[i].b) != cudaSuccess || cudaEventElapsedTime(&t, g_slots[i].a, g_slots[i].b) != cudaSuccess) {
            enf_set_error("timing_read: event query failed");
            return ENF_ERR_CUDA;
        }
        ms[g_slots[i].kind] += t;
        counts[g_slots[i].kind] += 1;
    }
    g_used = 0;
    return ENF_OK;
}

int64_t enflow_param_layout(int nf, int L, int64_t* offsets, int64_t* counts) {
    if (nf < 1 || nf > ENF_MAX_NF || L < 1 || L > 16) { enf_set_error("param_layout: bad nf=%d or L=%d", nf, L); return -1; }
    const EgclOffsets eo = enf_egcl_offsets(nf);
    const ArgmaxOffsets ao = enf_argmax_offsets(nf);
    int64_t esz[P_EGCL_COUNT], asz[PA_COUNT];
    enf_egcl_sizes(nf, esz);
    enf_argmax_sizes(nf, asz);
    // state_dict order differs from the internal enum order only in naming; list in egcl.py definition order
    static const int order[P_EGCL_COUNT] = {P_W1, P_B1, P_W2, P_B2, P_W4, P_B4, P_W5, P_B5, P_W3, P_B3, P_WC, P_W6, P_B6, P_W7, P_B7};
    int idx = 0;
    for (int l = 0; l < L; ++l)
        for (int t = 0; t < P_EGCL_COUNT; ++t, ++idx) {
            if (offsets) offsets[idx] = (int64_t)l * eo.size + eo.off[order[t]];
            if (counts) counts[idx] = esz[order[t]];
        }
    for (int t = 0; t < PA_COUNT; ++t, ++idx) {
        if (offsets) offsets[idx] = (int64_t)L * eo.size + ao.off[t];
        if (counts) counts[idx] = asz[t];
    }
    return (int64_t)L * eo.size + ao.size;
}

int enflow_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int* step,
                     float lr, const float* lr_dev, float beta1, float beta2, float eps, void* stream) {
    return enf_adam_step(params, grads, exp_avg, exp_avg_sq, n, step, lr, lr_dev, beta1, beta2, eps, ST(stream));
}

int64_t enflow_lj_prior_workspace_doubles(int N) { return enf_lj_prior_workspace_doubles(N); }
int enflow_lj_prior_forces(const double* pos, int N, const double* box, double softening, double cutoff, double* ws,
                           double* force, double* energy, void* stream) {
    ENF_CHECK_ARG(N >= 0 && box && box[0] > 0 && box[1] > 0 && box[2] > 0 && cutoff > 0, "lj_prior: bad box/cutoff");
    return enf_lj_prior_forces(pos, N, box, softening, cutoff, ws, force, energy, ST(stream));
}
int enflow_lj_prior_minimize(double* pos, int N, const double* box, double softening, double cutoff, int iters,
                             double rate, double cap, double* ws, void* stream) {
    ENF_CHECK_ARG(N >= 0 && box && box[0] > 0 && box[1] > 0 && box[2] > 0 && cutoff > 0, "lj_prior: bad box/cutoff");
    return enf_lj_prior_minimize(pos, N, box, softening, cutoff, iters, rate, cap, ws, ST(stream));
}
int enflow_lj_prior_velocities(double* vel, int N, double kBT, uint64_t seed, void* stream) {
    ENF_CHECK_ARG(N >= 0 && kBT >= 0, "lj_prior: negative size or temperature");
    return enf_lj_prior_velocities(vel, N, kBT, seed, ST(stream));
}
int enflow_lj_prior_run(double* pos, double* vel, int N, const double* box, double softening, double cutoff, double dt,
                        double a, double kBT, int n_steps, uint64_t seed, uint64_t step0, double* ws, double* energy,
                        void* stream) {
    ENF_CHECK_ARG(N >= 0 && box && box[0] > 0 && box[1] > 0 && box[2] > 0 && cutoff > 0, "lj_prior: bad box/cutoff");
    ENF_CHECK_ARG(a >= 0 && a <= 1 && kBT >= 0 && n_steps >= 0, "lj_prior: a=%g outside [0,1] or negative kBT/steps", a);
    return enf_lj_prior_run(pos, vel, N, box, softening, cutoff, dt, a, kBT, n_steps, seed, step0, ws, energy, ST(stream));
}

int64_t enflow_edges_workspace_ints(int N) { return enf_edges_workspace_ints(N); }

int enflow_build_edges(const void* pos, const void* box, int pos_is_f64, const float* r_cut, const int* mol_off,
                       int B, int N, int E_cap, int* row, int* col, int* rowptr, int* ref_pos, int* E_dev,
                       int* status, int* ws, void* stream) {
    ENF_CHECK_ARG(B >= 0 && N >= 0 && E_cap >= 0, "build_edges: negative size");
    if (pos_is_f64)
        return enf_build_edges_t<double>((const double*)pos, (const double*)box, r_cut, mol_off, B, N, E_cap, row, col,
                                         rowptr, ref_pos, E_dev, status, ws, ST(stream));
    return enf_build_edges_t<float>((const float*)pos, (const float*)box, r_cut, mol_off, B, N, E_cap, row, col,
                                    rowptr, ref_pos, E_dev, status, ws, ST(stream));
}

int enflow_fc_check(const float* pos, const float* box, const float* r_cut, const int* mol_off, int B, int* status,
                    void* stream) {
    return enf_fc_check(pos, box, r_cut, mol_off, B, status, ST(stream));
}
int enflow_fc_build(const int* mol_off, int B, int N, int E_cap, int* row, int* col, int* rowptr, int* E_dev,
                    int* colptr, int* perm, int* eoff, int* status, void* stream) {
    return enf_fc_build(mol_off, B, N, E_cap, row, col, rowptr, E_dev, colptr, perm, eoff, status, ST(stream));
}
int enflow_build_col_perm(const int* col, const int* rowptr, const int* mol_off, int B, int N, int E_cap,
                          const int* E_dev, int* colptr, int* perm, int* ws, void* stream) {
    return enf_build_col_perm(col, rowptr, mol_off, B, N, E_cap, E_dev, colptr, perm, ws, nullptr, ST(stream));
}

int enflow_segment_sum128(const float* x, const int* ptr, const int* perm, int N, int E_cap, int apply_silu,
                          float* out, void* stream) {
    return enf_segment_sum128(x, ptr, perm, N, E_cap, apply_silu, out, ST(stream));
}
int enflow_segment_sum3(const float* x, const int* ptr, const int* perm, int N, int E_cap, int mean, float scale,
                        int accumulate, float* out, void* stream) {
    return enf_segment_sum3(x, ptr, perm, N, E_cap, mean, scale, accumulate, out, ST(stream));
}

int64_t enflow_pack_floats(int nf) { return enf_pack_offsets(nf).size; }
int enflow_pack_layer(const float* lp, int nf, float* packed, void* stream) {
    return enf_pack_layer(lp, nf, packed, ST(stream));
}
int enflow_node_pre_fwd(const float* h, int N, int nf, const float* lp, float* P, float* S, float* Q, void* stream) {
    return enf_node_pre_fwd(h, N, nf, lp, P, S, Q, ST(stream));
}
int enflow_edge_fwd(const int* row, const int* col, const int* E_dev, int E_cap, const float* pos, const float* box,
                    const float* P, const float* S, const float* lp, const float* packed, int nf, float* wr,
                    float* z2, float* z3, float* s, float* trans, void* stream) {
    return enf_edge_fwd(row, col, E_dev, E_cap, pos, box, P, S, lp, packed, nf, wr, z2, z3, s, trans, ST(stream));
}
int64_t enflow_tc_pack_bytes(void) { return enf_tc_pack_bytes(); }
int enflow_tc_pack_layer(const float* lp, int nf, void* img, void* stream) {
    return enf_tc_pack_layer(lp, nf, (unsigned char*)img, ST(stream));
}
int enflow_edge_fwd_tc(int mode, const int* row, const int* col, const int* E_dev, int E_cap, const float* pos,
                       const float* box, const float* P, const float* S, const float* lp, const void* wimg, int nf,
                       const int* rowptr, const int* mis, float* runs, float* s, float* trans, void* stream) {
    ENF_CHECK_ARG(mode == 1 || mode == 2, "edge_fwd_tc: mode must be 1 (split) or 2 (bf16)");
    ENF_CHECK_ARG((reinterpret_cast<uintptr_t>(wimg) & 15) == 0, "edge_fwd_tc: weight image must be 16-byte aligned");
    return enf_edge_fwd_tc(mode, row, col, E_dev, E_cap, pos, box, P, S, lp, (const unsigned char*)wimg, nf, rowptr, mis,
                           runs, s, trans, ST(stream));
}
int64_t enflow_run_rows(int E_cap, int N) { return enf_run_rows(E_cap, N); }
int64_t enflow_run_scratch_ints(int N) { return enf_scan_scratch_ints(N + 2); }
int enflow_run_index(const int* rowptr, int N, int* mis, int* scratch, void* stream) {
    return enf_run_index(rowptr, N, mis, scratch, ST(stream));
}
int enflow_run_sum128(const float* runs, const int* rowptr, const int* mis, int N, int E_cap, float* out, void* stream) {
    return enf_run_sum128(runs, rowptr, mis, N, E_cap, out, ST(stream));
}
int enflow_node_post_fwd(const float* h, const float* agg, int N, int nf, const float* lp, const float* packed,
                         float* z4, float* G, void* stream) {
    return enf_node_post_fwd(h, agg, N, nf, lp, packed, z4, G, ST(stream));
}

int enflow_coupling_fwd(const float* Q, const float* F, const float* G, const float* h, const float* g,
                        const float* pos, const float* vel, const float* box, const int* mol_off, int B, int nf,
                        float dt, float* h_o, float* g_o, float* pos_o, float* vel_o, float* ldj_mol, void* stream) {
    return enf_coupling_fwd(Q, F, G, h, g, pos, vel, box, mol_off, B, nf, dt, h_o, g_o, pos_o, vel_o, ldj_mol, ST(stream));
}
int enflow_coupling_bwd(const float* Q, const float* vel_in, const float* dldj, int N, int nf, float dt, float* dh,
                        float* dg, float* dpos, float* dvel, float* dQ, float* dF, float* dG, void* stream) {
    return enf_coupling_bwd(Q, vel_in, dldj, N, nf, dt, dh, dg, dpos, dvel, dQ, dF, dG, ST(stream));
}
int enflow_coupling_inv_pre(const float* g, const float* vel, const float* box, int N, int nf, float dt, float* h,
                            float* pos, void* stream) {
    return enf_coupling_inv_pre(g, vel, box, N, nf, dt, h, pos, ST(stream));
}
int enflow_coupling_inv_post(const float* Q, const float* F, const float* G, const int* mol_off, int B, int nf,
                             float dt, float* g, float* vel, float* neg_ldj_mol, void* stream) {
    return enf_coupling_inv_post(Q, F, G, mol_off, B, nf, dt, g, vel, neg_ldj_mol, ST(stream));
}

int enflow_argmax_fwd(const float* h, const float* eps, int N, int nf, const float* ap, const int* mol_off, int B,
                      float* z, float* logq_atom, double* logq_mol, float* log_q, void* stream) {
    return enf_argmax_fwd(h, eps, N, nf, ap, mol_off, B, z, logq_atom, logq_mol, log_q, ST(stream));
}

int enflow_nll_slices(int max_n) { return enf_nll_slices(max_n); }

int enflow_nll_fwd(const float* pos, const float* vel, const float* h, const float* g, const int* mol_off, int B,
                   int N, int nf, int max_n, float kBT, float softening, float z_lj, const float* ldj,
                   double* mol_term, float* loss, void* stream) {
    enf_time_begin(TK_NLL, ST(stream));
    const int rc = enf_nll_fwd(pos, vel, h, g, mol_off, B, N, nf, max_n, kBT, softening, z_lj, ldj, mol_term, loss, ST(stream));
    enf_time_end(ST(stream));
    return rc;
}
int enflow_nll_bwd(const float* pos, const float* vel, const float* h, const float* g, const int* mol_off, int B,
                   int nf, int max_n, float kBT, float softening, const float* dloss, float* dpos, float* dvel,
                   float* dh, float* dg, float* dldj, void* stream) {
    enf_time_begin(TK_NLL, ST(stream));
    const int rc = enf_nll_bwd(pos, vel, h, g, mol_off, B, nf, max_n, kBT, softening, dloss, dpos, dvel, dh, dg, dldj, ST(stream));
    enf_time_end(ST(stream));
    return rc;
}

}  // extern "C"
#pragma GCC visibility pop
