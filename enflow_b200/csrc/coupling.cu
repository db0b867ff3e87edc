// K3: the leap-frog coupling step of LFIntegrator (enflow/flow/dynamics.py:14-21) fused into one
// pass over the node state, with log|det J| accumulated per molecule by a fixed-order warp reduction:
//   vel' = exp(Q) vel + F dt ; g' = g + G dt ; pos' = wrap(pos + vel' dt, box) ; h' = h + g' dt ; ldj_mol += sum_atoms Q
// (ldj uses sum(Q), not 3 sum(Q): quirk Q2.)  Backward and the exact inverse (dynamics.py:25-37) follow.
// HBM-bound: algorithmic bytes per atom = (19 + 5 nf) * 4, + 4 B per molecule.
#include "common.cuh"

namespace {

// One CTA per group of mols_per_cta consecutive molecules.  Their atoms are contiguous, so the [*,3] and [*,nf]
// arrays are walked as flat element ranges (every load/store fully coalesced, all 32 lanes busy whatever the
// molecule size); the per-molecule sum of Q is then a warp reduction (one warp per molecule, fixed xor tree).

__global__ void __launch_bounds__(256) k_coupling_fwd(const float* __restrict__ Q, const float* __restrict__ F,
                                                       const float* __restrict__ G, const float* __restrict__ h,
                                                       const float* __restrict__ g, const float* __restrict__ pos,
                                                       const float* __restrict__ vel, const float* __restrict__ box,
                                                       const int* __restrict__ mol_off, int B, int nf, float dt,
                                                       int mols_per_cta, float* __restrict__ h_o,
                                                       float* __restrict__ g_o, float* __restrict__ pos_o,
                                                       float* __restrict__ vel_o, float* __restrict__ ldj_mol) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int m0 = blockIdx.x * mols_per_cta; m0 < B; m0 += gridDim.x * mols_per_cta) {
        const int m1 = min(m0 + mols_per_cta, B);
        const int a0 = mol_off[m0], a1 = mol_off[m1];
        const int64_t v0 = 3LL * a0;
        const int nv = 3 * (a1 - a0);
        for (int t = threadIdx.x; t < nv; t += 256) {                         // vel, pos (dynamics.py:14,17-18)
            const int64_t idx = v0 + t;
            const float v = fmaf(expf(Q[a0 + t / 3]), vel[idx], F[idx] * dt);
            vel_o[idx] = v;
            pos_o[idx] = wrapf_(fmaf(v, dt, pos[idx]), box[idx]);
        }
        const int64_t f0 = (int64_t)nf * a0;
        const int nfe = nf * (a1 - a0);
        for (int t = threadIdx.x; t < nfe; t += 256) {                        // g, h (dynamics.py:15,19)
            const int64_t idx = f0 + t;
            const float gn = fmaf(G[idx], dt, g[idx]);
            g_o[idx] = gn;
            h_o[idx] = fmaf(gn, dt, h[idx]);
        }
        for (int m = m0 + wid; m < m1; m += 8) {                              // ldj += sum Q (dynamics.py:21, Q2)
            float qs = 0.f;
            for (int i = mol_off[m] + lane; i < mol_off[m + 1]; i += 32) qs += Q[i];
            qs = warp_sum(qs);
            if (lane == 0) ldj_mol[m] += qs;
        }
    }
}

// thread per atom. In/out gradient buffers hold d/d(outputs) on entry and d/d(inputs) on exit
// (EGCL backward then adds its own dh / dpos contributions).
__global__ void __launch_bounds__(256) k_coupling_bwd(const float* __restrict__ Q, const float* __restrict__ vel_in,
                                                       const float* __restrict__ dldj, int N, int nf, float dt,
                                                       float* __restrict__ dh, float* __restrict__ dg,
                                                       float* __restrict__ dpos, float* __restrict__ dvel,
                                                       float* __restrict__ dQ, float* __restrict__ dF,
                                                       float* __restrict__ dG) {
    const float dl = dldj[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float s = expf(Q[i]);
        float dq = dl;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float dvt = fmaf(dpos[(int64_t)i * 3 + c], dt, dvel[(int64_t)i * 3 + c]);
            dF[(int64_t)i * 3 + c] = dvt * dt;
            dq = fmaf(dvt * vel_in[(int64_t)i * 3 + c], s, dq);
            dvel[(int64_t)i * 3 + c] = dvt * s;
        }
        dQ[i] = dq;
        for (int c = 0; c < nf; ++c) {
            const float dgt = fmaf(dh[(int64_t)i * nf + c], dt, dg[(int64_t)i * nf + c]);
            dg[(int64_t)i * nf + c] = dgt;
            dG[(int64_t)i * nf + c] = dgt * dt;
        }
    }
}

// inverse, first half (dynamics.py:27-29): h -= g dt ; pos = wrap(pos - vel dt)
__global__ void __launch_bounds__(256) k_coupling_inv_pre(const float* __restrict__ g, const float* __restrict__ vel,
                                                           const float* __restrict__ box, int N, int nf, float dt,
                                                           float* __restrict__ h, float* __restrict__ pos) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            pos[(int64_t)i * 3 + c] =
                wrapf_(fmaf(-vel[(int64_t)i * 3 + c], dt, pos[(int64_t)i * 3 + c]), box[(int64_t)i * 3 + c]);
        for (int c = 0; c < nf; ++c) h[(int64_t)i * nf + c] = fmaf(-g[(int64_t)i * nf + c], dt, h[(int64_t)i * nf + c]);
    }
}

// inverse, second half (dynamics.py:32-33): g -= G dt ; vel = (vel - F dt)/exp(Q) ; neg_ldj_mol -= sum Q
__global__ void __launch_bounds__(256) k_coupling_inv_post(const float* __restrict__ Q, const float* __restrict__ F,
                                                            const float* __restrict__ G,
                                                            const int* __restrict__ mol_off, int B, int nf, float dt,
                                                            float* __restrict__ g, float* __restrict__ vel,
                                                            float* __restrict__ neg_ldj_mol) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < B; m += warps) {
        const int a0 = mol_off[m], a1 = mol_off[m + 1];
        float qs = 0.f;
        for (int i = a0 + lane; i < a1; i += 32) {
            const float q = Q[i];
            qs += q;
            const float inv = expf(-q);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                vel[(int64_t)i * 3 + c] = fmaf(-F[(int64_t)i * 3 + c], dt, vel[(int64_t)i * 3 + c]) * inv;
            for (int c = 0; c < nf; ++c) g[(int64_t)i * nf + c] = fmaf(-G[(int64_t)i * nf + c], dt, g[(int64_t)i * nf + c]);
        }
        qs = warp_sum(qs);
        if (lane == 0 && neg_ldj_mol) neg_ldj_mol[m] -= qs;
    }
}

}  // namespace

static int mol_grid(int B) {
    int blocks = (B + 7) / 8;
    int cap = enf_num_sms() * 8;
    return blocks < cap ? (blocks > 0 ? blocks : 1) : cap;
}
static int atom_grid(int N) {
    int blocks = (N + 255) / 256;
    int cap = enf_num_sms() * 8;
    return blocks < cap ? (blocks > 0 ? blocks : 1) : cap;
}

int enf_coupling_fwd(const float* Q, const float* F, const float* G, const float* h, const float* g,
                     const float* pos, const float* vel, const float* box, const int* mol_off, int B, int nf,
                     float dt, float* h_o, float* g_o, float* pos_o, float* vel_o, float* ldj_mol, cudaStream_t st) {
    if (B == 0) return ENF_OK;
    // enough CTAs to fill the machine twice at small batches, up to 32 molecules per CTA at large ones
    int mpc = B / (2 * enf_num_sms());
    mpc = mpc < 1 ? 1 : (mpc > 32 ? 32 : mpc);
    const int cgrid = (B + mpc - 1) / mpc;
    enf_count_launch(), k_coupling_fwd<<<cgrid, 256, 0, st>>>(Q, F, G, h, g, pos, vel, box, mol_off, B, nf, dt, mpc, h_o, g_o,
                                                              pos_o, vel_o, ldj_mol);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

int enf_coupling_bwd(const float* Q, const float* vel_in, const float* dldj, int N, int nf, float dt, float* dh,
                     float* dg, float* dpos, float* dvel, float* dQ, float* dF, float* dG, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    enf_count_launch(), k_coupling_bwd<<<atom_grid(N), 256, 0, st>>>(Q, vel_in, dldj, N, nf, dt, dh, dg, dpos, dvel, dQ, dF, dG);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

int enf_coupling_inv_pre(const float* g, const float* vel, const float* box, int N, int nf, float dt, float* h,
                         float* pos, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    enf_count_launch(), k_coupling_inv_pre<<<atom_grid(N), 256, 0, st>>>(g, vel, box, N, nf, dt, h, pos);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

int enf_coupling_inv_post(const float* Q, const float* F, const float* G, const int* mol_off, int B, int nf,
                          float dt, float* g, float* vel, float* neg_ldj_mol, cudaStream_t st) {
    if (B == 0) return ENF_OK;
    enf_count_launch(), k_coupling_inv_post<<<mol_grid(B), 256, 0, st>>>(Q, F, G, mol_off, B, nf, dt, g, vel, neg_ldj_mol);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}
