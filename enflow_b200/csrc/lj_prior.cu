// Prior sampler for generate: Langevin dynamics of the periodic soft Lennard-Jones fluid that the
// reference runs through OpenMM (enflow/data/lj.py:32-89, enflow/data/simulated.py:83-132):
//   U = sum_{i<j, r<rc} 4 ((1/(s + r))^12 - (1/(s + r))^6)      (lj.py:65, sigma = eps = 1 in reduced units, s = softening,
//                                                                 CutoffPeriodic with minimum image, no switching, lj.py:74-75)
//   LangevinMiddleIntegrator (simulated.py:109):  v += dt f ; x += dt/2 v ; v = a v + sqrt(kT (1 - a^2)) R ; x += dt/2 v
// in the unit system of the likelihood (mass 1, enflow/flow/loss.py:16-22), so the frames are samples of the
// density Alchemical_NLL assumes for (pos, vel).  State and forces are fp64: one system of a few thousand atoms,
// thousands of steps, nothing here is throughput-critical, and drift-free energies make the sampler testable.
//
// Forces: thread = atom, the partner range split over gridDim.y slabs, slab partials added in slab order
// (deterministic).  Noise: Philox4x32-10 keyed by the seed, counter = (step, atom): reproducible and stateless.
#include "common.cuh"

namespace {

constexpr int FT = 128;          // threads per force CTA

__device__ __forceinline__ double min_image(double d, double L) { return d - L * rint(d / L); }

__global__ void __launch_bounds__(FT) k_lj_forces(const double* __restrict__ pos, int N, double bx, double by, double bz,
                                                   double soft, double rc, double* __restrict__ fpart,
                                                   double* __restrict__ upart) {
    __shared__ double sx[FT], sy[FT], sz[FT];
    const int i = blockIdx.x * FT + threadIdx.x;
    const int slabs = gridDim.y;
    const int per = (N + slabs - 1) / slabs;
    const int j0 = blockIdx.y * per, j1 = min(N, j0 + per);
    double xi = 0.0, yi = 0.0, zi = 0.0;
    if (i < N) { xi = pos[3 * i]; yi = pos[3 * i + 1]; zi = pos[3 * i + 2]; }
    double fx = 0.0, fy = 0.0, fz = 0.0, u = 0.0;
    const double rc2 = rc * rc;
    for (int t0 = j0; t0 < j1; t0 += FT) {
        const int j = t0 + threadIdx.x;
        __syncthreads();
        if (j < j1) { sx[threadIdx.x] = pos[3 * j]; sy[threadIdx.x] = pos[3 * j + 1]; sz[threadIdx.x] = pos[3 * j + 2]; }
        __syncthreads();
        const int cnt = min(FT, j1 - t0);
        if (i < N) {
            for (int k = 0; k < cnt; ++k) {
                if (t0 + k == i) continue;
                const double dx = min_image(xi - sx[k], bx), dy = min_image(yi - sy[k], by), dz = min_image(zi - sz[k], bz);
                const double r2 = dx * dx + dy * dy + dz * dz;
                if (r2 < rc2) {
                    const double r = sqrt(r2);
                    const double q = 1.0 / (soft + r);
                    const double q2 = q * q, q6 = q2 * q2 * q2, q12 = q6 * q6;
                    u += 4.0 * (q12 - q6);
                    // -dU/dr = 4 (12 q^13 - 6 q^7); force on i along d / r
                    const double g = 4.0 * (12.0 * q12 - 6.0 * q6) * q / fmax(r, 1e-300);
                    fx += g * dx; fy += g * dy; fz += g * dz;
                }
            }
        }
    }
    if (i < N) {
        double* f = fpart + ((int64_t)blockIdx.y * N + i) * 3;
        f[0] = fx; f[1] = fy; f[2] = fz;
        upart[(int64_t)blockIdx.y * N + i] = 0.5 * u;        // each pair is seen from both ends
    }
}

__global__ void k_lj_combine(const double* __restrict__ fpart, const double* __restrict__ upart, int N, int slabs,
                             double* __restrict__ force, double* __restrict__ u_atom) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double fx = 0.0, fy = 0.0, fz = 0.0, u = 0.0;
    for (int s = 0; s < slabs; ++s) {
        const double* f = fpart + ((int64_t)s * N + i) * 3;
        fx += f[0]; fy += f[1]; fz += f[2];
        u += upart[(int64_t)s * N + i];
    }
    force[3 * i] = fx; force[3 * i + 1] = fy; force[3 * i + 2] = fz;
    if (u_atom) u_atom[i] = u;
}

// ---- Philox4x32-10 (Salmon et al. 2011): counter-based, one call gives four 32-bit words
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
// three standard normals for (stream, step, atom): Box-Muller on 53-bit-free uniforms in (0, 1]
__device__ __forceinline__ void normal3(uint64_t seed, uint32_t stream, uint64_t step, uint32_t atom, double (&n)[3]) {
    uint32_t c[4] = {atom, stream, (uint32_t)step, (uint32_t)(step >> 32)};
    philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u0 = ((double)c[0] + 1.0) * (1.0 / 4294967296.0), u1 = (double)c[1] * (1.0 / 4294967296.0);
    const double u2 = ((double)c[2] + 1.0) * (1.0 / 4294967296.0), u3 = (double)c[3] * (1.0 / 4294967296.0);
    const double r0 = sqrt(-2.0 * log(u0)), r1 = sqrt(-2.0 * log(u2));
    double s0, c0, s1, c1;
    sincospi(2.0 * u1, &s0, &c0);
    sincospi(2.0 * u3, &s1, &c1);
    n[0] = r0 * c0; n[1] = r0 * s0; n[2] = r1 * c1;
    (void)s1;
}

// LangevinMiddle step, everything after the force evaluation
__global__ void k_langevin_middle(double* __restrict__ pos, double* __restrict__ vel, const double* __restrict__ force,
                                  int N, double dt, double a, double b, uint64_t seed, uint64_t step) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double R[3];
    normal3(seed, 0u, step, (uint32_t)i, R);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double v = vel[3 * i + c] + dt * force[3 * i + c];
        double x = pos[3 * i + c] + 0.5 * dt * v;
        v = a * v + b * R[c];
        x += 0.5 * dt * v;
        vel[3 * i + c] = v;
        pos[3 * i + c] = x;
    }
}

// steepest descent with a capped displacement (stands in for OpenMM's minimizeEnergy, simulated.py:113)
__global__ void k_descend(double* __restrict__ pos, const double* __restrict__ force, int N, double rate, double cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * N) return;
    pos[i] += fmax(-cap, fmin(cap, rate * force[i]));
}

__global__ void k_maxwell(double* __restrict__ vel, int N, double sd, uint64_t seed, uint32_t stream) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double R[3];
    normal3(seed, stream, 0, (uint32_t)i, R);
    vel[3 * i] = sd * R[0]; vel[3 * i + 1] = sd * R[1]; vel[3 * i + 2] = sd * R[2];
}

__global__ void k_energy(const double* __restrict__ u_atom, const double* __restrict__ vel, int N, double* __restrict__ out) {
    __shared__ double su[256], sk[256];
    double u = 0.0, k = 0.0;
    for (int i = threadIdx.x; i < N; i += 256) {
        u += u_atom[i];
        if (vel) k += 0.5 * (vel[3 * i] * vel[3 * i] + vel[3 * i + 1] * vel[3 * i + 1] + vel[3 * i + 2] * vel[3 * i + 2]);
    }
    su[threadIdx.x] = u; sk[threadIdx.x] = k;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { su[threadIdx.x] += su[threadIdx.x + s]; sk[threadIdx.x] += sk[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = su[0]; out[1] = sk[0]; }
}

int slabs_for(int N) {
    int s = (4 * enf_num_sms() * FT) / (N > 0 ? N : 1);      // ~4 CTAs per SM
    return s < 1 ? 1 : (s > 32 ? 32 : s);
}

int forces(const double* pos, int N, const double* box, double soft, double rc, double* ws, double* force,
           double* u_atom, cudaStream_t st) {
    const int slabs = slabs_for(N);
    double* fpart = ws;
    double* upart = ws + (size_t)slabs * N * 3;
    enf_count_launch(), k_lj_forces<<<dim3((N + FT - 1) / FT, slabs), FT, 0, st>>>(pos, N, box[0], box[1], box[2], soft, rc, fpart, upart);
    enf_count_launch(), k_lj_combine<<<(N + 255) / 256, 256, 0, st>>>(fpart, upart, N, slabs, force, u_atom);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

}  // namespace

// workspace (doubles): slab partials (4 per atom per slab) + force [3N] + u_atom [N]
int64_t enf_lj_prior_workspace_doubles(int N) { return (int64_t)slabs_for(N) * N * 4 + 4LL * N; }

int enf_lj_prior_forces(const double* pos, int N, const double* box, double soft, double rc, double* ws, double* force,
                        double* energy, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    double* u_atom = ws + (size_t)slabs_for(N) * N * 4 + 3LL * N;
    ENF_TRY(forces(pos, N, box, soft, rc, ws, force, u_atom, st));
    if (energy) {          // {potential, 0}
        enf_count_launch(), k_energy<<<1, 256, 0, st>>>(u_atom, nullptr, N, energy);
        ENF_CHECK_LAUNCH();
    }
    return ENF_OK;
}

int enf_lj_prior_minimize(double* pos, int N, const double* box, double soft, double rc, int iters, double rate,
                          double cap, double* ws, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    double* force = ws + (size_t)slabs_for(N) * N * 4;
    for (int it = 0; it < iters; ++it) {
        ENF_TRY(forces(pos, N, box, soft, rc, ws, force, nullptr, st));
        enf_count_launch(), k_descend<<<(3 * N + 255) / 256, 256, 0, st>>>(pos, force, N, rate, cap);
    }
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

int enf_lj_prior_velocities(double* vel, int N, double kBT, uint64_t seed, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    enf_count_launch(), k_maxwell<<<(N + 255) / 256, 256, 0, st>>>(vel, N, sqrt(kBT), seed, 1u);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

// n_steps LangevinMiddle steps starting at step counter step0; energy (optional) = {potential, kinetic} after the last step
int enf_lj_prior_run(double* pos, double* vel, int N, const double* box, double soft, double rc, double dt, double a,
                     double kBT, int n_steps, uint64_t seed, uint64_t step0, double* ws, double* energy,
                     cudaStream_t st) {
    if (N == 0) return ENF_OK;
    double* force = ws + (size_t)slabs_for(N) * N * 4;
    double* u_atom = force + 3LL * N;
    const double b = sqrt(kBT * (1.0 - a * a));
    for (int s = 0; s < n_steps; ++s) {
        ENF_TRY(forces(pos, N, box, soft, rc, ws, force, nullptr, st));
        enf_count_launch(), k_langevin_middle<<<(N + 255) / 256, 256, 0, st>>>(pos, vel, force, N, dt, a, b, seed, step0 + s);
    }
    if (energy) {
        ENF_TRY(forces(pos, N, box, soft, rc, ws, force, u_atom, st));
        enf_count_launch(), k_energy<<<1, 256, 0, st>>>(u_atom, vel, N, energy);
    }
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}
