// Secular-equation root finder shared by the warp-per-root CUDA kernel (secular_kernel in
// cuppen_kernels.cu) and the host-side unit tests (tests/ compile this header with g++ to
// check the numerics without a GPU).  It replaces the bisection loop of the reference,
// /root/reference/src/eigenvalues.c:161-247 (secularEquation :8-17), with a rational
// interpolation ("middle way", Li 1994 / LAPACK working note 89) iteration that is safeguarded
// by a bracket, and returns the root as (origin pole, tau) so that consumers can form
// d_j - lambda_i = (d_j - d_origin) - tau without cancellation.
//
// Canonical problem (the caller reflects rho<0 problems):  rho > 0,
//   d[0] < d[1] < ... < d[k-1],  w[j] = z_j^2 > 0,
//   g(lambda) = 1/rho + sum_j w[j] / (d[j] - lambda),   root i in (d[i], d[i+1]),
//   last root in (d[k-1], d[k-1] + rho*sum(w)].
#ifndef CUPPEN_SECULAR_CORE_H
#define CUPPEN_SECULAR_CORE_H

#include <math.h>

#ifndef CUPPEN_RCP
#define CUPPEN_RCP(x) (1.0 / (x))      // (platform.h supplies the device version: hardware seed + one cubic step)
#endif
#ifndef CUPPEN_HD
#if defined(__CUDACC__)
#define CUPPEN_HD __host__ __device__ __forceinline__
#else
#define CUPPEN_HD inline
#endif
#endif

namespace cuppen {

struct SecularSums {
    double psi, dpsi, phi, dphi, err;
};

#ifndef CUPPEN_PLATFORM_H
// Lanes policy for the host: one lane, no reduction (platform.h has the full version).
struct SerialLanes {
    CUPPEN_HD int lane() const { return 0; }
    CUPPEN_HD int lanes() const { return 1; }
    CUPPEN_HD double sum(double v) const { return v; }
};
#endif

// psi = sum_{j<=split} w_j/(delta_j - tau), phi = sum_{j>split}; derivatives likewise.
// Branch-free loop bodies, unrolled: four independent reciprocal chains per lane keep the FP64
// pipe busy.
template <class Lanes>
CUPPEN_HD SecularSums secular_eval(const Lanes& L, int k, const double* __restrict__ d,
                                   const double* __restrict__ w, double dorg, double tau, int split) {
    // (every term of psi is negative and every term of phi positive -- the poles up to `split` lie below the root, the
    // others above --, so the sum of the |terms| that the stopping criterion needs is phi - psi: no tenth FP64
    // instruction per pole for it)
    double psi = 0, dpsi = 0, phi = 0, dphi = 0;
    const int nl = L.lanes();
    int j = L.lane();
    const int split1 = split + 1 < k ? split + 1 : k;
#pragma unroll 4
    for (; j < split1; j += nl) {
        const double t = (d[j] - dorg) - tau;
        const double inv = CUPPEN_RCP(t);
        const double r = w[j] * inv;
        psi += r; dpsi += r * inv;
    }
#pragma unroll 4
    for (; j < k; j += nl) {
        const double t = (d[j] - dorg) - tau;
        const double inv = CUPPEN_RCP(t);
        const double r = w[j] * inv;
        phi += r; dphi += r * inv;
    }
    SecularSums s;
    s.psi = L.sum(psi); s.dpsi = L.sum(dpsi); s.phi = L.sum(phi); s.dphi = L.sum(dphi);
    s.err = fabs(s.psi) + fabs(s.phi);
    return s;
}

struct SecularRoot {
    int origin;     // index of the pole tau is measured from
    double tau;     // lambda = d[origin] + tau
    int iters;      // function evaluations spent
};

// The default evaluator: every call streams the k poles through the lanes of one warp (or one host thread).
template <class Lanes>
struct SecularStreamEval {
    const Lanes& L;
    int k;
    const double* d;
    const double* w;
    CUPPEN_HD SecularSums operator()(double dorg, double tau, int split) const { return secular_eval(L, k, d, w, dorg, tau, split); }
};

// Solve for root i.  sumw = sum_j w[j] (only needed for the last root).  `ev(dorg, tau, split)` evaluates the
// sums over all k poles (secular_kernel passes a CTA-collective evaluator that stages the poles in shared memory).
template <class Eval>
CUPPEN_HD SecularRoot secular_solve_ev(const Eval& ev, int k, const double* __restrict__ d,
                                       const double* __restrict__ w, double rho, double sumw, int i) {
    const double eps = 2.220446049250313e-16;
    const double rhoinv = 1.0 / rho;
    SecularRoot out;
    out.iters = 0;
    if (k == 1) { out.origin = 0; out.tau = rho * w[0]; return out; }

    const bool last = (i == k - 1);
    // interpolation poles: (ip0, ip1) = (i, i+1) for interior roots, (k-2, k-1) for the last one
    const int ip0 = last ? k - 2 : i;
    const int ip1 = ip0 + 1;
    int org;
    double lo, hi, tau;
    const double gap = d[ip1] - d[ip0];

    if (!last) {
        // decide the origin from the sign of g at the midpoint (evaluated relative to pole i)
        const double half = 0.5 * gap;
        SecularSums s = ev(d[i], half, i);
        out.iters++;
        const double gmid = rhoinv + s.psi + s.phi;              // full secular function at the midpoint
        const double c = gmid - w[ip0] / (-half) - w[ip1] / half;   // without the two nearest poles (initial guess only)
        const double a_ = c * gap, w0 = w[ip0], w1 = w[ip1];
        if (gmid >= 0.0) {            // root in the left half: origin i, tau in (0, gap/2]
            org = i; lo = 0.0; hi = half;
            const double a = a_ + w0 + w1, b = w0 * gap;
            const double disc = sqrt(fabs(a * a - 4.0 * b * c));
            tau = (a > 0.0) ? 2.0 * b / (a + disc) : (a - disc) / (2.0 * c);
        } else {                      // origin i+1, tau in [-gap/2, 0)
            org = i + 1; lo = -half; hi = 0.0;
            const double a = a_ - w0 - w1, b = w1 * gap;
            const double disc = sqrt(fabs(a * a + 4.0 * b * c));
            tau = (a < 0.0) ? 2.0 * b / (a - disc) : -(a + disc) / (2.0 * c);
        }
        if (!(tau > lo && tau < hi)) tau = 0.5 * (lo + hi);
    } else {
        org = k - 1;
        const double R = rho * sumw;
        const double mid = 0.5 * R;
        SecularSums s = ev(d[org], mid, k);
        out.iters++;
        const double w0 = w[ip0], w1 = w[ip1];
        const double gmid = rhoinv + s.psi;                      // full secular function at the midpoint
        const double c = gmid - w0 / (-gap - mid) - w1 / (-mid); // poles 0..k-3 only (initial guess)
        const double a = -c * gap + w0 + w1, b = w1 * gap;
        const double disc = sqrt(fabs(a * a + 4.0 * b * c));
        if (gmid <= 0.0) {            // root above the midpoint
            lo = mid; hi = R;
            const double temp = w0 / (gap + R) + w1 / R;
            if (c <= temp) tau = R;
            else tau = (a < 0.0) ? 2.0 * b / (disc - a) : (a + disc) / (2.0 * c);
        } else {
            lo = 0.0; hi = mid;
            tau = (a < 0.0) ? 2.0 * b / (disc - a) : (a + disc) / (2.0 * c);
        }
        if (!(tau > lo && tau <= hi)) tau = 0.5 * (lo + hi);
    }

    const double dorg = d[org];
    const int split = last ? k - 2 : i;     // psi covers poles <= split
    double prevabs = INFINITY;
    int slow = 0;
    for (int it = 0; it < 80; ++it) {
        SecularSums s = ev(dorg, tau, split);
        out.iters++;
        const double h = rhoinv + s.psi + s.phi;
        const double errb = eps * (8.0 * s.err + fabs(rhoinv) + fabs(tau) * (s.dpsi + s.dphi));
        if (fabs(h) <= errb) break;
        if (h < 0.0) lo = tau; else hi = tau;                    // g is increasing in tau
        if (hi - lo <= 2.0 * eps * fmax(fabs(lo), fabs(hi))) { tau = 0.5 * (lo + hi); break; }
        // middle-way step: match psi, psi' with a pole at ip0 and phi, phi' with a pole at ip1
        const double DL = (d[ip0] - dorg) - tau, DR = (d[ip1] - dorg) - tau;
        const double c = h - DL * s.dpsi - DR * s.dphi;
        const double a = (DL + DR) * h - DL * DR * (s.dpsi + s.dphi);
        const double b = DL * DR * h;
        double eta;
        if (c == 0.0) {
            eta = (a != 0.0) ? b / a : 0.5 * (lo + hi) - tau;
        } else {
            const double disc = sqrt(fabs(a * a - 4.0 * b * c));
            if (!last) eta = (a <= 0.0) ? (a - disc) / (2.0 * c) : 2.0 * b / (a + disc);
            else       eta = (a >= 0.0) ? (a + disc) / (2.0 * c) : 2.0 * b / (a - disc);
        }
        if (h * eta >= 0.0) eta = -h / (s.dpsi + s.dphi);        // wrong direction: Newton step
        double nt = tau + eta;
        const double ah = fabs(h);
        if (ah > 0.5 * prevabs) slow++;
        prevabs = ah;
        if (!(nt > lo && nt < hi) || slow >= 2) {
            // safeguard: bisect the bracket (geometrically when it spans many binades)
            if (lo > 0.0 && hi > 16.0 * lo) nt = sqrt(lo) * sqrt(hi);
            else if (hi < 0.0 && lo < 16.0 * hi) nt = -sqrt(-lo) * sqrt(-hi);
            else nt = 0.5 * (lo + hi);
            slow = 0;
        }
        tau = nt;
    }
    out.origin = org;
    out.tau = tau;
    return out;
}

template <class Lanes>
CUPPEN_HD SecularRoot secular_solve(const Lanes& L, int k, const double* __restrict__ d,
                                    const double* __restrict__ w, double rho, double sumw, int i) {
    return secular_solve_ev(SecularStreamEval<Lanes>{L, k, d, w}, k, d, w, rho, sumw, i);
}

}  // namespace cuppen
#endif
