// Exact first-passage sampler: the rejection method of Tuerlinckx, Maris, Ratcliff & De Boeck
// (2001) as the reference runs it in pyhddmjagsutils.py:47-176 (`simulratcliff`, the data generator of its JAGS / Stan
// scripts, alpha_not_scaled.py:95-97).  No time step: a trial is a chain of exits from the largest interval that is
// symmetric around the current position and fits between the boundaries; each exit time is drawn by rejection from
// the series (16) of the paper.  One thread per trial, fp64 (these generators make 1e4 trials, not 1e9), uniforms
// from the trial's own Philox stream (stream 3), so results do not depend on the launch shape.
//
// Kept from the reference: the per-trial drift Nu + Eta*z with |Nu| clipped to 5 and Eta == 0 replaced by 1e-16
// (:104-111; its `eta = 3` clip assigns a name that is never read, so Eta is NOT clipped there, nor here), the start
// point and non-decision time ranges, a fresh non-decision time draw on every interval (:160), the `delta = eps`
// boundary test (:161-168) and the signed-RT output (:173).  Different: the choice probability is evaluated as
// 1 / (1 + exp(-x)) (the reference's exp(x) / (1 + exp(x)) is NaN beyond x = 709), mu == 0 and a start on a boundary
// are given their limits instead of NaNs, and loops are capped.
#include "ddm_kernels.cuh"

namespace ddm {

constexpr uint32_t STREAM_EXACT = 3u;

// Sequential uniforms in (0, 1) with 52 random bits each, two per Philox block of the trial's stream.
struct UniformStream {
    uint32_t block, trial, dataset, have;
    const PhiloxKey &key;
    uint32_t w[4];
    __device__ UniformStream(uint32_t trial_, uint32_t dataset_, const PhiloxKey &key_)
        : block(0u), trial(trial_), dataset(dataset_), have(0u), key(key_) {}
    __device__ double next() {
        if (have == 0u) {
            philox4x32_rk(block++, trial, dataset, STREAM_EXACT, key, w);
            have = 2u;
        }
        const uint32_t hi = w[2u * (2u - have)], lo = w[2u * (2u - have) + 1u];
        have--;
        const uint64_t bits = ((uint64_t)(hi >> 6) << 26) | (uint64_t)(lo >> 6);  // 52 bits
        return ((double)bits + 0.5) * (1.0 / 4503599627370496.0);
    }
};

struct ExactArgs {
    const double *params;  // [n_datasets * 8]: Alpha, Tau, Nu, Beta, rangeTau, rangeBeta, Eta, Varsigma
    double *out;           // [n_datasets * n_trials] signed response times
    unsigned long long *stats;
    uint32_t n_datasets, n_trials, dataset_offset, trial_offset;
    PhiloxKey key;
};

__global__ void __launch_bounds__(128) exact_sampler_kernel(const ExactArgs a, uint64_t total) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const uint32_t ds = (uint32_t)(g / a.n_trials);
    const uint32_t trial = (uint32_t)(g - (uint64_t)ds * a.n_trials);
    const double *p = a.params + (size_t)ds * 8;
    const double Alpha = p[0], Tau = p[1], rangeTau = p[4], rangeBeta = p[5], Varsigma = p[7];
    double Nu = p[2], Eta = p[6];
    if (Nu < -5.0 || Nu > 5.0) Nu = copysign(5.0, Nu);
    if (Eta == 0.0) Eta = 1e-16;
    const double D = Varsigma * Varsigma / 2.0;
    const double eps = 2.220446049250313e-16, delta = eps;
    const double pi = 3.14159265358979323846;

    UniformStream u(trial + a.trial_offset, ds + a.dataset_offset, a.key);
    const double u1 = u.next(), u2 = u.next();
    const double r1 = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    const double mu = Nu + r1 * Eta;
    const double bb = p[3] - rangeBeta / 2.0 + rangeBeta * u.next();
    const double zz = bb * Alpha;
    const double Aupper = Alpha - zz, Alower = -zz;
    double radius = fmin(fabs(Aupper), fabs(Alower));
    double totaltime = 0.0, startpos = 0.0, T = 0.0, X = 0.0;
    unsigned long long intervals = 0;
    bool capped = false;
    for (int it = 0;; it++) {
        const double ndt_now = Tau - rangeTau / 2.0;
        if (!(radius > 0.0) || it >= 100000) {  // on a boundary already (or no progress): absorbed where it stands
            capped = it >= 100000;
            T = ndt_now + rangeTau * u.next() + totaltime;
            X = (fabs(Aupper - startpos) <= fabs(Alower - startpos)) ? 1.0 : -1.0;
            break;
        }
        intervals++;
        const double lambda = 0.25 * mu * mu / D + 0.25 * D * pi * pi / (radius * radius);
        double F = 1.0;  // limit of F^2 / (1 + F^2) for mu -> 0
        if (mu != 0.0) {
            const double f = D * pi / (radius * mu);
            const double f2 = f * f;
            F = isinf(f2) ? 1.0 : f2 / (1.0 + f2);
        }
        const double prob = 1.0 / (1.0 + exp(-radius * mu / D));
        const double dir = (u.next() < prob) ? 1.0 : -1.0;
        double l = -1.0, s2 = 0.0, s1 = 0.5;
        for (int rej = 0; s2 > l && rej < 10000; rej++) {
            s2 = u.next();
            s1 = u.next();
            double tnew = 0.0, told = 0.0;
            const double ls1 = log(s1);
            for (int uu = 1; uu <= 200; uu++) {
                told = tnew;
                const double k = 2.0 * uu + 1.0;
                const double term = k * exp(F * k * k * ls1);  // (2uu+1) * s1^(F (2uu+1)^2)
                tnew = told + ((uu & 1) ? -term : term);
                if (!(fabs(tnew - told) > eps)) break;
            }
            l = 1.0 + exp(-F * ls1) * tnew;  // 1 + s1^(-F) * sum
        }
        const double t = fabs(log(s1)) / lambda;
        totaltime += t;
        const double pos = startpos + dir * radius;
        const double ndt = ndt_now + rangeTau * u.next();
        if (pos + delta > Aupper) {
            T = ndt + totaltime;
            X = 1.0;
            break;
        } else if (pos - delta < Alower) {
            T = ndt + totaltime;
            X = -1.0;
            break;
        }
        startpos = pos;
        radius = fmin(fabs(Aupper - startpos), fabs(Alower - startpos));
    }
    a.out[g] = T * X;
    atomicAdd(a.stats + STAT_STEPS, intervals);
    if (X > 0.0) atomicAdd(a.stats + STAT_UPPER, 1ull);
    if (capped) atomicAdd(a.stats + STAT_REJECT_CAP, 1ull);
}

cudaError_t launch_exact_sampler(const double *params, double *out, unsigned long long *stats, uint32_t n_datasets,
                                 uint32_t n_trials, uint32_t dataset_offset, uint32_t trial_offset, const PhiloxKey &key,
                                 cudaStream_t s) {
    const uint64_t total = (uint64_t)n_datasets * n_trials;
    if (total == 0) return cudaSuccess;
    ExactArgs a{params, out, stats, n_datasets, n_trials, dataset_offset, trial_offset, key};
    exact_sampler_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(a, total);
    return cudaGetLastError();
}

}  // namespace ddm
