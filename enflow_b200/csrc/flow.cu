// Whole-pass orchestration: LFIntegrator.forward, its backward, and LFIntegrator.reverse
// (enflow/flow/dynamics.py:10-37) as one stream-ordered sequence of kernel launches per call.
// No allocation and no host synchronisation happen here: the caller provides one workspace that is
// carved deterministically from the dims, and data-dependent sizes (edge counts) stay on the device.
#include "internal.h"

struct enflow_dims_t {
    int32_t B, N, nf, L, E_cap, max_n;
    float dt, coords_weight;
    int32_t mode;      // 0 = fp32 on the FFMA pipe, 1 = tcgen05 bf16x3 split (fp32-accurate), 2 = tcgen05 bf16
    int32_t fc;        // 1 = fully connected regime assumed: implicit all-pairs list built once, re-proved per layer (fc.cu)
};

#define TIMED(kind, call)        \
    do {                         \
        enf_time_begin(kind, st); \
        ENF_TRY(call);           \
        enf_time_end(st);        \
    } while (0)

namespace {

// side streams are used whenever the per-family timing (eager, one event pair per kernel family on the caller's stream) is off
bool side_ok() { return !enf_timing_on(); }

struct Bump {
    char* base;
    size_t off;
    explicit Bump(void* p) : base(reinterpret_cast<char*>(p)), off(0) {}
    template <typename T>
    T* take(size_t count) {
        off = (off + 255) / 256 * 256;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct LayerSave {           // kept per layer when training
    int *row, *col, *rowptr, *E_dev, *mis;
    float *Q, *z2, *z3, *s, *agg, *z4;     // z2/z3 only in mode 0 (the tensor-core backward recomputes them)
    float *P, *S;                          // node-level halves of edge_nn.0 (2 x N x H: cheaper to keep than to redo)
};

struct Workspace {
    // states[l] = node state entering layer l; states[L] only when training is off is the output itself
    float *h[17], *g[17], *pos[17], *vel[17];
    LayerSave layer[16];
    float* packed;           // L * pack size
    unsigned char* tcimg;    // L * swizzled bf16 weight images for the tcgen05 kernels
    float *F, *G, *trans, *wr, *runs;
    int* run_scratch;
    int* edges_ws;
    int* fc_eoff;            // per-molecule edge offsets of the implicit all-pairs list (fc mode)
    float* logq_atom;
    double* logq_mol;
    float* log_q;
    // backward scratch
    float *dQ, *dF, *dG, *dagg, *dP, *dS, *dz1, *dd;
    float *partial, *partial_post, *partial_pre;      // per-CTA weight-gradient partials: edge kernel (and argmax), node_post, node_pre
    unsigned char* geom;     // per-tile edge records of the tensor-core backward kernel
    int *colptr, *perm, *same;
    size_t bytes;
};

size_t partial_floats(const enflow_dims_t& d) {
    size_t m = (size_t)enf_edge_partial_floats();
    size_t a = (size_t)enf_node_post_partial_floats(d.N, d.nf);
    if (a > m) m = a;
    a = (size_t)enf_node_pre_partial_floats(d.N, d.nf);
    if (a > m) m = a;
    a = (size_t)enf_argmax_partial_floats(d.N, d.nf);
    if (a > m) m = a;
    return m;
}

Workspace carve(const enflow_dims_t& d, void* base, int training) {
    Workspace w;
    Bump b(base);
    const size_t N = d.N, nf = d.nf, E = d.E_cap, H = ENF_H;
    const int saved_layers = training ? d.L : 1;
    const int nstates = training ? d.L : 2;        // inference ping-pongs between two buffers
    for (int l = 0; l < nstates; ++l) {
        w.h[l] = b.take<float>(N * nf); w.g[l] = b.take<float>(N * nf);
        w.pos[l] = b.take<float>(N * 3); w.vel[l] = b.take<float>(N * 3);
    }
    for (int l = 0; l < saved_layers; ++l) {
        LayerSave& s = w.layer[l];
        if (d.fc && l > 0) {          // one neighbour list for every coupling step
            const LayerSave& s0 = w.layer[0];
            s.row = s0.row; s.col = s0.col; s.rowptr = s0.rowptr; s.E_dev = s0.E_dev; s.mis = s0.mis;
        } else {
            s.row = b.take<int>(E); s.col = b.take<int>(E); s.rowptr = b.take<int>(N + 1); s.E_dev = b.take<int>(2);
            s.mis = b.take<int>(N + 2);
        }
        s.Q = b.take<float>(N); s.s = b.take<float>(E);
        s.z2 = d.mode == 0 ? b.take<float>(E * H) : nullptr;
        s.z3 = d.mode == 0 ? b.take<float>(E * H) : nullptr;
        s.agg = b.take<float>(N * H); s.z4 = b.take<float>(N * H);
        s.P = b.take<float>(N * H); s.S = b.take<float>(N * H);
    }
    w.packed = b.take<float>((size_t)d.L * enf_pack_offsets(d.nf).size);
    w.tcimg = b.take<unsigned char>((size_t)d.L * enf_tc_pack_bytes() + 1024);
    w.F = b.take<float>(N * 3); w.G = b.take<float>(N * nf);
    w.trans = b.take<float>(E * 3); w.wr = b.take<float>(H);
    w.runs = d.mode ? b.take<float>((size_t)enf_run_rows(d.E_cap, d.N) * H) : nullptr;
    w.run_scratch = b.take<int>((size_t)enf_scan_scratch_ints(d.N + 2));
    w.edges_ws = b.take<int>((size_t)enf_edges_workspace_ints(d.N));
    w.fc_eoff = b.take<int>((size_t)d.B + 2);
    w.logq_atom = b.take<float>(N); w.logq_mol = b.take<double>(d.B); w.log_q = b.take<float>(1);
    if (training) {
        w.dQ = b.take<float>(N); w.dF = b.take<float>(N * 3); w.dG = b.take<float>(N * nf);
        w.dagg = b.take<float>(N * H); w.dP = b.take<float>(N * H); w.dS = b.take<float>(N * H);
        w.dz1 = b.take<float>(E * H); w.dd = b.take<float>(E * 3);
        w.partial = b.take<float>(partial_floats(d));
        w.partial_post = b.take<float>((size_t)enf_node_post_partial_floats(d.N, d.nf));
        w.partial_pre = b.take<float>((size_t)enf_node_pre_partial_floats(d.N, d.nf));
        w.geom = d.mode ? b.take<unsigned char>((size_t)enf_edge_bwd_geom_bytes(d.E_cap)) : nullptr;
        w.colptr = b.take<int>(N + 1); w.perm = b.take<int>(E); w.same = b.take<int>(4);
    }
    w.bytes = (b.off + 255) / 256 * 256;
    return w;
}

int check_dims(const enflow_dims_t* d) {
    ENF_CHECK_ARG(d != nullptr, "dims is NULL");
    ENF_CHECK_ARG(d->nf >= 1 && d->nf <= ENF_MAX_NF, "nf=%d outside [1,%d]", d->nf, ENF_MAX_NF);
    ENF_CHECK_ARG(d->L >= 1 && d->L <= 16, "L=%d outside [1,16]", d->L);
    ENF_CHECK_ARG(d->B >= 0 && d->N >= 0 && d->E_cap >= 0, "negative size");
    ENF_CHECK_ARG(d->mode >= 0 && d->mode <= 2, "mode=%d outside [0,2]", d->mode);
    ENF_CHECK_ARG(d->fc == 0 || d->fc == 1, "fc=%d is not 0 or 1", d->fc);
    return ENF_OK;
}

const float* layer_params(const float* params, int nf, int l) { return params + (int64_t)l * enf_egcl_offsets(nf).size; }
float* layer_params(float* params, int nf, int l) { return params + (int64_t)l * enf_egcl_offsets(nf).size; }
const float* argmax_params(const float* params, int nf, int L) { return params + (int64_t)L * enf_egcl_offsets(nf).size; }
float* argmax_params(float* params, int nf, int L) { return params + (int64_t)L * enf_egcl_offsets(nf).size; }

unsigned char* tc_image(const Workspace& w, int l) {      // 1024-byte aligned: bulk copies need 16
    unsigned char* base = w.tcimg + ((1024 - (reinterpret_cast<uintptr_t>(w.tcimg) & 1023)) & 1023);
    return base + (int64_t)l * enf_tc_pack_bytes();
}

void copy_f(float* dst, const float* src, size_t n, cudaStream_t st) {
    if (dst != src && n) cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
}

// Q, F, G of one EGCL at the given node state (egcl.py:77-93); saves activations into `sv`
int egcl_forward(const enflow_dims_t& d, const Workspace& w, const LayerSave& sv, const float* lp, const float* packed,
                 const unsigned char* tcimg, const float* h, const float* pos, const float* box, const float* r_cut, const int* mol_off,
                 int* status, bool keep, cudaStream_t st) {
    if (d.fc)      // the list exists (fc_lists); prove that these positions are still in the regime where it is the answer
        TIMED(TK_EDGES, enf_fc_check(pos, box, r_cut, mol_off, d.B, status, st));
    else
        TIMED(TK_EDGES, enf_build_edges_t<float>(pos, box, r_cut, mol_off, d.B, d.N, d.E_cap, sv.row, sv.col, sv.rowptr, nullptr,
                                         sv.E_dev, status, w.edges_ws, st));
    TIMED(TK_NODE_PRE, enf_node_pre_fwd(h, d.N, d.nf, lp, sv.P, sv.S, sv.Q, st));
    if (d.mode == 0) {
        TIMED(TK_EDGE_FWD, enf_edge_fwd(sv.row, sv.col, sv.E_dev, d.E_cap, pos, box, sv.P, sv.S, lp, packed, d.nf, w.wr,
                                        sv.z2, sv.z3, sv.s, w.trans, st));
        TIMED(TK_SEG_ROWS, enf_segment_sum128(sv.z2, sv.rowptr, nullptr, d.N, d.E_cap, 1, sv.agg, st));
    } else {
        // tensor-core path: the kernel reduces silu(z2) over each row into per-run partials; no [E,H] store
        if (!d.fc) ENF_TRY(enf_run_index(sv.rowptr, d.N, sv.mis, w.run_scratch, st));
        TIMED(TK_EDGE_FWD, enf_edge_fwd_tc(d.mode, sv.row, sv.col, sv.E_dev, d.E_cap, pos, box, sv.P, sv.S, lp, tcimg, d.nf,
                                           sv.rowptr, sv.mis, w.runs, sv.s, w.trans, st));
        // agg = row sums of the run partials, F = coords_weight * row means of trans (helpers.py:62-70): one launch
        TIMED(TK_RUN_SUM, enf_run_sum128_sum3(w.runs, sv.rowptr, sv.mis, d.N, d.E_cap, sv.agg, w.trans, 1, d.coords_weight, 0,
                                             w.F, st));
    }
    if (d.mode == 0)
        TIMED(TK_SEG3, enf_segment_sum3(w.trans, sv.rowptr, nullptr, d.N, d.E_cap, 1, d.coords_weight, 0, w.F, st));
    if (d.mode == 0)
        TIMED(TK_NODE_POST, enf_node_post_fwd(h, sv.agg, d.N, d.nf, lp, packed, sv.z4, w.G, st));
    else
        TIMED(TK_NODE_POST, enf_node_post_fwd_tc(d.mode, h, sv.agg, d.N, d.nf, lp, tcimg, keep ? sv.z4 : nullptr, w.G, st));      // z4: backward only
    return ENF_OK;
}

// fc mode: the all-pairs list, its run index and (training) its column-grouped permutation, once per pass
int fc_lists(const enflow_dims_t& d, const Workspace& w, const int* mol_off, int training, int* status, cudaStream_t st) {
    const LayerSave& sv = w.layer[0];
    TIMED(TK_EDGES, enf_fc_build(mol_off, d.B, d.N, d.E_cap, sv.row, sv.col, sv.rowptr, sv.E_dev, training ? w.colptr : nullptr,
                                 training ? w.perm : nullptr, w.fc_eoff, status, st));
    if (d.mode) ENF_TRY(enf_run_index(sv.rowptr, d.N, sv.mis, w.run_scratch, st));
    return ENF_OK;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" size_t enflow_flow_workspace_bytes(const enflow_dims_t* dims, int training) {
    if (check_dims(dims) != ENF_OK) return 0;
    return carve(*dims, nullptr, training).bytes;
}

extern "C" int enflow_flow_forward(const enflow_dims_t* dims, const float* params, const float* h_in,
                                   const float* g_in, const float* pos_in, const float* vel_in, const float* box,
                                   const float* r_cut, const int* mol_off, const float* eps, void* workspace,
                                   size_t workspace_bytes, int training, float* h_out, float* g_out, float* pos_out,
                                   float* vel_out, float* ldj_mol, float* ldj, int* status, void* stream) {
    ENF_TRY(check_dims(dims));
    const enflow_dims_t& d = *dims;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Workspace w = carve(d, workspace, training);
    ENF_CHECK_ARG(workspace_bytes >= w.bytes, "workspace too small: %zu < %zu", workspace_bytes, w.bytes);
    const int nf = d.nf;
    const size_t N = d.N;
    if (d.B == 0 || d.N == 0) {          // empty batch: nothing to transform, log-det 0
        cudaMemsetAsync(ldj, 0, sizeof(float), st);
        return ENF_OK;
    }
    cudaMemsetAsync(ldj_mol, 0, sizeof(float) * d.B, st);
    // weight images: the FFMA kernels' transposed fp32 copies (mode 0) or the swizzled bf16 hi/lo images of all
    // layers in two launches (tensor-core modes)
    if (d.mode == 0) {
        for (int l = 0; l < d.L; ++l)
            ENF_TRY(enf_pack_layer(layer_params(params, nf, l), nf, w.packed + (int64_t)l * enf_pack_offsets(nf).size, st));
    } else {
        ENF_TRY(enf_tc_pack_layers(params, nf, d.L, enf_egcl_offsets(nf).size, tc_image(w, 0), enf_tc_pack_bytes(), st));
    }
    // state entering layer 0: dequantised h (dynamics.py:11), everything else copied
    if (eps) {
        TIMED(TK_ARGMAX, enf_argmax_fwd(h_in, eps, d.N, nf, argmax_params(params, nf, d.L), mol_off, d.B, w.h[0], w.logq_atom,
                               w.logq_mol, w.log_q, st));
    } else {
        copy_f(w.h[0], h_in, N * nf, st);
    }
    copy_f(w.g[0], g_in, N * nf, st);
    copy_f(w.pos[0], pos_in, N * 3, st);
    copy_f(w.vel[0], vel_in, N * 3, st);
    if (d.fc) ENF_TRY(fc_lists(d, w, mol_off, training, status, st));
    for (int l = 0; l < d.L; ++l) {
        const int cur = training ? l : (l & 1);
        const bool last = l == d.L - 1;
        const int nxt = training ? l + 1 : ((l + 1) & 1);
        float* ho = last ? h_out : w.h[nxt];
        float* go = last ? g_out : w.g[nxt];
        float* po = last ? pos_out : w.pos[nxt];
        float* vo = last ? vel_out : w.vel[nxt];
        const LayerSave& sv = w.layer[training ? l : 0];
        ENF_TRY(egcl_forward(d, w, sv, layer_params(params, nf, l), w.packed + (int64_t)l * enf_pack_offsets(nf).size,
                             tc_image(w, l), w.h[cur], w.pos[cur], box, r_cut, mol_off, status, training != 0, st));
        TIMED(TK_COUPLING_FWD, enf_coupling_fwd(sv.Q, w.F, w.G, w.h[cur], w.g[cur], w.pos[cur], w.vel[cur], box, mol_off, d.B, nf, d.dt,
                                 ho, go, po, vo, ldj_mol, st));
    }
    ENF_TRY(enf_ldj_total(ldj_mol, d.B, eps ? w.log_q : nullptr, ldj, st));
    return ENF_OK;
}

extern "C" int enflow_flow_backward(const enflow_dims_t* dims, const float* params, float* grads, const float* h_in,
                                    const float* box, const int* mol_off, const float* eps, void* workspace,
                                    size_t workspace_bytes, float* dh, float* dg, float* dpos, float* dvel,
                                    const float* dldj, int* status, void* stream) {
    ENF_TRY(check_dims(dims));
    const enflow_dims_t& d = *dims;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Workspace w = carve(d, workspace, 1);
    ENF_CHECK_ARG(workspace_bytes >= w.bytes, "workspace too small: %zu < %zu", workspace_bytes, w.bytes);
    const int nf = d.nf;
    if (d.B == 0 || d.N == 0) return ENF_OK;
    // Tensor-core modes: the reductions of the per-CTA weight-gradient partials (three per layer, nothing reads their result
    // before the optimizer) run on a side stream, in order, beside the kernels of the main chain (C2: 6.04 -> 5.96 ms per
    // step).  Each producer has its own partial buffer and waits for the reduction that last read it (one layer earlier:
    // long finished).  (Also tried: the per-tile edge records and the regime check on a second side stream: no gain.)
    const bool par = d.mode != 0 && side_ok();
    cudaStream_t red = par ? enf_side_stream(0) : st;
    enf_chain(st, red);                               // the gradient buffer is ready on st (zeroed by the caller)
    for (int l = d.L - 1; l >= 0; --l) {
        const LayerSave& sv = w.layer[l];
        const float* lp = layer_params(params, nf, l);
        float* lg = layer_params(grads, nf, l);
        // coupling step (dynamics.py:14-21): gradients w.r.t. Q, F, G and the incoming state
        TIMED(TK_COUPLING_BWD, enf_coupling_bwd(sv.Q, w.vel[l], dldj, d.N, nf, d.dt, dh, dg, dpos, dvel, w.dQ, w.dF, w.dG, st));
        // node_model (egcl.py:65-69)
        if (d.mode == 0)
            TIMED(TK_NODE_POST_BWD, enf_node_post_bwd(w.h[l], sv.agg, sv.z4, w.dG, d.N, nf, lp,
                                                 w.packed + (int64_t)l * enf_pack_offsets(nf).size, w.dagg, dh, lg, w.partial, st));
        else {
            if (par && l < d.L - 1) enf_wait_mark(0, st);      // (layer l + 1's reduction has read partial_post: long ago)
            TIMED(TK_NODE_POST_BWD, enf_node_post_bwd_tc(d.mode, w.h[l], sv.agg, sv.z4, w.dG, d.N, nf, lp, tc_image(w, l), w.dagg,
                                                    dh, lg, w.partial_post, st, red));
            if (par) enf_mark(0, red);
        }
        // edge_model + force_model (egcl.py:57-63,71-75); P/S were kept by the forward pass
        // column-grouped view of the edges; reused as is when this layer's list equals the one just processed
        // (fully connected regime: the neighbour list is the same at every coupling step)
        if (!d.fc) {      // (fc mode: the forward pass left the permutation of the one list in colptr / perm)
            const int* skip = nullptr;
            if (l < d.L - 1) {
                const LayerSave& nx = w.layer[l + 1];
                ENF_TRY(enf_edges_same(sv.row, sv.col, sv.E_dev, nx.row, nx.col, nx.E_dev, w.same, st));
                skip = w.same;
            }
            TIMED(TK_COL_PERM, enf_build_col_perm(sv.col, sv.rowptr, mol_off, d.B, d.N, d.E_cap, sv.E_dev, w.colptr, w.perm,
                                                  w.edges_ws, skip, st));
        }
        if (d.mode == 0) {
            TIMED(TK_EDGE_BWD, enf_edge_bwd(sv.row, sv.col, sv.rowptr, sv.E_dev, d.E_cap, w.pos[l], box, sv.P, sv.S, lp, nf,
                                            w.wr, sv.z2, sv.z3, sv.s, w.dagg, w.dF, d.coords_weight, w.dz1, w.dd, lg,
                                            w.partial, st));
            TIMED(TK_SEG_ROWS, enf_segment_sum128(w.dz1, sv.rowptr, nullptr, d.N, d.E_cap, 0, w.dP, st));
        } else {
            ENF_TRY(enf_edge_bwd_tc_geom(sv.row, sv.col, sv.rowptr, sv.E_dev, d.E_cap, w.pos[l], box, sv.s, w.dF, d.coords_weight,
                                         sv.mis, w.geom, st));
            if (par && l < d.L - 1) enf_wait_mark(1, st);
            ENF_TRY(enf_edge_bwd_tc(d.mode, sv.E_dev, d.E_cap, sv.P, sv.S, lp, tc_image(w, l), nf, w.dagg, w.runs, w.dz1, w.dd, lg,
                                    w.partial, w.geom, status, st, red));      // (times its kernels itself)
            if (par) enf_mark(1, red);
            // dP = the rows' run sums: on the side stream (behind this layer's edge reduction), beside the column sum below
            TIMED(TK_RUN_SUM, enf_run_sum128(w.runs, sv.rowptr, sv.mis, d.N, d.E_cap, w.dP, red));
            if (par) enf_mark(3, red);
        }
        if (d.mode == 0) TIMED(TK_SEG3, enf_segment_sum3(w.dd, sv.rowptr, nullptr, d.N, d.E_cap, 0, 1.0f, 1, dpos, st));
        // dS through the column permutation, and d loss / d pos: +dd onto row atoms (coord_diff = pos[row] - pos[col],
        // data/base.py:17), then -dd onto col atoms through the permutation (tensor-core modes: both in this launch)
        TIMED(TK_SEG_COLS, enf_segment_sum128_sum3(w.dz1, w.colptr, w.perm, d.N, d.E_cap, 0, w.dS, w.dd, -1.0f, dpos,
                                                   d.mode != 0 ? sv.rowptr : nullptr, st));
        if (par) enf_wait_mark(3, st);                       // dP is complete
        if (par && l < d.L - 1) enf_wait_mark(2, st);
        TIMED(TK_NODE_PRE_BWD, enf_node_pre_bwd(w.h[l], d.N, nf, lp, w.dP, w.dS, w.dQ, dh, lg, d.mode != 0 ? w.partial_pre : w.partial, st,
                                                 d.mode != 0 ? red : st));
        if (par) enf_mark(2, red);
    }
    enf_chain(red, st);                               // every weight gradient is in place before the caller goes on
    if (eps)
        TIMED(TK_ARGMAX, enf_argmax_bwd(h_in, eps, d.N, nf, argmax_params(params, nf, d.L), dh, dldj,
                               argmax_params(grads, nf, d.L), w.partial, st));
    return ENF_OK;
}

extern "C" int enflow_flow_reverse(const enflow_dims_t* dims, const float* params, float* h, float* g, float* pos,
                                   float* vel, const float* box, const float* r_cut, const int* mol_off,
                                   void* workspace, size_t workspace_bytes, int quantize, float* neg_ldj_mol,
                                   int* status, void* stream) {
    ENF_TRY(check_dims(dims));
    const enflow_dims_t& d = *dims;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Workspace w = carve(d, workspace, 0);
    ENF_CHECK_ARG(workspace_bytes >= w.bytes, "workspace too small: %zu < %zu", workspace_bytes, w.bytes);
    const int nf = d.nf;
    if (d.B == 0 || d.N == 0) return ENF_OK;
    if (neg_ldj_mol) cudaMemsetAsync(neg_ldj_mol, 0, sizeof(float) * d.B, st);
    // weight images: the FFMA kernels' transposed fp32 copies (mode 0) or the swizzled bf16 hi/lo images of all
    // layers in two launches (tensor-core modes)
    if (d.mode == 0) {
        for (int l = 0; l < d.L; ++l)
            ENF_TRY(enf_pack_layer(layer_params(params, nf, l), nf, w.packed + (int64_t)l * enf_pack_offsets(nf).size, st));
    } else {
        ENF_TRY(enf_tc_pack_layers(params, nf, d.L, enf_egcl_offsets(nf).size, tc_image(w, 0), enf_tc_pack_bytes(), st));
    }
    const LayerSave& sv = w.layer[0];
    if (d.fc) ENF_TRY(fc_lists(d, w, mol_off, 0, status, st));
    for (int l = d.L - 1; l >= 0; --l) {
        TIMED(TK_COUPLING_INV, enf_coupling_inv_pre(g, vel, box, d.N, nf, d.dt, h, pos, st));                    // dynamics.py:27-29
        ENF_TRY(egcl_forward(d, w, sv, layer_params(params, nf, l), w.packed + (int64_t)l * enf_pack_offsets(nf).size,
                             tc_image(w, l), h, pos, box, r_cut, mol_off, status, false, st));                    // :31
        TIMED(TK_COUPLING_INV, enf_coupling_inv_post(sv.Q, w.F, w.G, mol_off, d.B, nf, d.dt, g, vel, neg_ldj_mol, st));   // :32-33
    }
    if (quantize) ENF_TRY(enf_argmax_reverse(h, d.N, nf, st));                                    // :35
    return ENF_OK;
}
#pragma GCC visibility pop
