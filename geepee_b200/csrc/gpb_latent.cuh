// gpb_latent.cuh -- elementwise latent-variable kernels: the
// Gaussian natural-parameter algebra per (row, latent dimension) that surrounds the sparse-GP layer
// in the latent-variable models.  One thread per (row, q); block-reduced sums for the log-partition
// terms (two-stage, deterministic).
//   a12 SGPLVM: get_cavity_x aep_models.py:840-861, compute_phi_x 863-867, compute_cav_grad_x 817-838,
//               get_posterior_x base_models.py:765-775, compute_posterior_grad_x 913-929;
//               VFE twin vfe_models.py:749-845 (KL of q(x), 857-863)
//   a13 SGPSSM: compute_cavity_x aep_models.py:1376-1387, compute_transition_tilted 1317-1374 (2-D branch),
//               compute_posterior_grad_x 1208-1232, compute_logZ_grad_x 1234-1285, compute_cavity_grad_x
//               1287-1315, compute_phi_{posterior,cavity}_x 1389-1437, posterior tying base_models.py:1719-1725
#pragma once

namespace gpb {

struct LvmArgs {
    int mode;            // 0 = AEP (cavity of x), 1 = VFE (posterior of x)
    int nat;             // natural parameters (x1 = precision*mean factor, x2 = log sqrt precision factor)
    const double* x1;    // [N, Q] raw parameter
    const double* x2;    // [N, Q] raw parameter (log of the square root)
    const long* sel;     // [n] row indices, or NULL: rows lo .. lo + n - 1
    long lo;
    int n, Q;
    double prior1, prior2;   // AEP: prior natural parameters; VFE: prior mean m0 and variance v0
    double alpha, s_cav, s_post;   // VFE: s_cav = sx = N / n
    const double* dmx;   // [n, Q] gradients from the layer (backward only)
    const double* dvx;
    double* o1;          // fwd: m [n, Q]     bwd: gx1 [N, Q] (rows `sel` written)
    double* o2;          // fwd: v [n, Q]     bwd: gx2 [N, Q]
    double* part;        // bwd: [gridDim.x][2] block sums {phi_cav | klx, phi_post | 0}
};

GPB_DEVICE void lvm_moments(const LvmArgs& a, double f1, double f2, double& m, double& v, double& p1, double& p2) {
    // posterior naturals (base_models.py:899-911) and the propagated moments
    if (a.nat) {
        p1 = (a.mode == 0 ? a.prior1 : a.prior1 / a.prior2) + f1;
        p2 = (a.mode == 0 ? a.prior2 : 1.0 / a.prior2) + f2;
    } else {
        p1 = f1 / f2;
        p2 = 1.0 / f2;
    }
    if (a.mode == 1) {
        m = p1 / p2;
        v = 1.0 / p2;
        return;
    }
    double c1, c2;
    if (a.nat) {
        c1 = a.prior1 + (1.0 - a.alpha) * f1;
        c2 = a.prior2 + (1.0 - a.alpha) * f2;
    } else {
        c1 = a.prior1 + (f1 / f2 - a.prior1) * (1.0 - a.alpha);
        c2 = a.prior2 + (1.0 / f2 - a.prior2) * (1.0 - a.alpha);
    }
    m = c1 / c2;
    v = 1.0 / c2;
}

GPB_KERNEL void lvm_x_fwd_kernel(LvmArgs a) {
    const long total = (long)a.n * a.Q;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const long i = idx / a.Q;
        const int q = (int)(idx % a.Q);
        const long row = a.sel ? a.sel[i] : a.lo + i;
        const double f1 = a.x1[row * a.Q + q], f2 = exp(2.0 * a.x2[row * a.Q + q]);
        double m, v, p1, p2;
        lvm_moments(a, f1, f2, m, v, p1, p2);
        a.o1[idx] = m;
        a.o2[idx] = v;
    }
}

GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) lvm_x_bwd_kernel(LvmArgs a) {
    GPB_SHARED double scratch[8];
    const long total = (long)a.n * a.Q;
    double s0 = 0.0, s1 = 0.0;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const long i = idx / a.Q;
        const int q = (int)(idx % a.Q);
        const long row = a.sel ? a.sel[i] : a.lo + i;
        const double f1 = a.x1[row * a.Q + q], f2 = exp(2.0 * a.x2[row * a.Q + q]);
        double m, v, p1, p2;
        lvm_moments(a, f1, f2, m, v, p1, p2);
        double d1, d2;
        if (a.mode == 1) {
            // KL(q(x) || p(x)) and its chain rule (vfe_models.py:857-863, base_models.py:913-929)
            const double m0 = a.prior1, v0 = a.prior2;
            s0 += 0.5 * (log(v0) - log(v) + (v + (m - m0) * (m - m0)) / v0 - 1.0);
            const double dm = a.dmx[idx] + a.s_cav * (m - m0) / v0;
            const double dv = a.dvx[idx] + a.s_cav * (-0.5 / v + 0.5 / v0);
            if (a.nat) {
                d1 = dm / p2;
                d2 = (-dm * p1 / (p2 * p2) - dv / (p2 * p2)) * 2.0 * f2;
            } else {
                d1 = dm;
                d2 = dv * 2.0 * f2;
            }
        } else {
            const double mpost = p1 / p2, vpost = 1.0 / p2;
            // phi_x of the cavity and of the posterior (aep_models.py:863-867) and their derivatives
            s0 += 0.5 * (m * m / v + log(v));
            s1 += 0.5 * (mpost * mpost / vpost + log(vpost));
            const double dmc = a.s_cav * (m / v) + a.dmx[idx];
            const double dvc = a.s_cav * 0.5 * (-m * m / (v * v) + 1.0 / v) + a.dvx[idx];
            const double dmp = a.s_post * (mpost / vpost);
            const double dvp = a.s_post * 0.5 * (-mpost * mpost / (vpost * vpost) + 1.0 / vpost);
            const double t1 = m / v, t2 = 1.0 / v;
            d1 = (1.0 - a.alpha) * dmc / t2;
            d2 = (1.0 - a.alpha) * (-dmc * t1 / (t2 * t2) - dvc / (t2 * t2));
            if (a.nat) {
                d2 = d2 * 2.0 * f2;
                d1 = d1 + dmp / p2;
                d2 = d2 + (-dmp * p1 / (p2 * p2) - dvp / (p2 * p2)) * 2.0 * f2;
            } else {
                const double dmq = d1 / f2;
                const double dvq = -d1 * f1 / (f2 * f2) - d2 / (f2 * f2);
                d1 = dmq + dmp;
                d2 = (dvq + dvp) * 2.0 * f2;
            }
        }
        a.o1[row * a.Q + q] = d1;
        a.o2[row * a.Q + q] = d2;
    }
    const double r0 = block_sum(s0, scratch);
    const double r1 = block_sum(s1, scratch);
    if (threadIdx.x == 0) {
        a.part[2 * blockIdx.x] = r0;
        a.part[2 * blockIdx.x + 1] = r1;
    }
}

// ---- SGPSSM --------------------------------------------------------------------------------
struct SsmArgs {
    const double* xf1;   // [T, Q] x_factor_1
    const double* xf2;   // [T, Q] x_factor_2 (log of the square root)
    long T;
    int Q;
    double prior1, prior2, alpha;
};
// posterior / cavity naturals of latent state t (base_models.py:1719-1725, aep_models.py:1385-1387)
GPB_DEVICE void ssm_naturals(const SsmArgs& a, long t, int q, double& f1, double& f2, double& p1, double& p2,
                             double& c1, double& c2, double& w) {
    f1 = a.xf1[t * a.Q + q];
    f2 = exp(2.0 * a.xf2[t * a.Q + q]);
    w = (t == 0 || t == a.T - 1) ? 2.0 : 3.0;
    p1 = w * f1;
    p2 = w * f2;
    if (t == 0) {
        p1 += a.prior1;
        p2 += a.prior2;
    }
    c1 = p1 - a.alpha * f1;
    c2 = p2 - a.alpha * f2;
}

GPB_KERNEL void ssm_cavity_kernel(SsmArgs a, double* __restrict__ cav_m, double* __restrict__ cav_v) {
    const long total = a.T * a.Q;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        double f1, f2, p1, p2, c1, c2, w;
        ssm_naturals(a, idx / a.Q, (int)(idx % a.Q), f1, f2, p1, p2, c1, c2, w);
        cav_m[idx] = c1 / (c2 + 1e-16);
        cav_v[idx] = 1.0 / (c2 + 1e-16);
    }
}

// transition factor t -> t+1, 2-D branch (aep_models.py:1334-1348): targets = cavity of state t+1
//   out: dm_layer = -dmt (what the layer's backward takes), dvt; part[block] = {sum lz, sum dvt}
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) ssm_transition_kernel(
    const double* __restrict__ mt, const double* __restrict__ vt, const double* __restrict__ mp,
    const double* __restrict__ vp, const double* __restrict__ sn, long total, double alpha, double s_dyn,
    double* __restrict__ dm_layer, double* __restrict__ dvt, double* __restrict__ part) {
    GPB_SHARED double scratch[8];
    const double sn2 = exp(2.0 * sn[0]);
    double s0 = 0.0, s1 = 0.0;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const double vsum = vt[idx] + vp[idx] + sn2 / alpha;
        const double md = mt[idx] - mp[idx];
        s0 += -0.5 * md * md / vsum - 0.5 * log(1.0 + alpha * (vt[idx] + vp[idx]) / sn2)
            - 0.5 * alpha * log(2.0 * 3.14159265358979323846 * sn2);
        const double dv = s_dyn * (-0.5 / vsum + 0.5 * md * md / (vsum * vsum));
        dvt[idx] = dv;
        dm_layer[idx] = s_dyn * (md / vsum);
        s1 += dv;
    }
    const double r0 = block_sum(s0, scratch);
    const double r1 = block_sum(s1, scratch);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = r0;
        part[2 * blockIdx.x + 1] = r1;
    }
}

// l1 / l2 [T, Q]: the three logZ sources of every latent state chained to its cavity naturals
// (aep_models.py:1252-1275): "prev" (target of the transition from t-1: rows p0+1 .. p0+np of dmt = -dm_layer,
// dvt), "next" (input of the transition from t: rows n0 .. n0+nn-1 of the layer's dmx, dvx) and "up" (input of the
// emission at t: rows u0 .. u0+nu-1).  Rows without a source get 0.
struct SsmSrc {
    const double* dm;    // [count, ld]
    const double* dv;
    long first, count;   // rows first .. first + count - 1 of the series
    int ld;
    double sign;         // -1 for "prev" (dm holds -dmt)
};
GPB_KERNEL void ssm_sources_kernel(SsmArgs a, SsmSrc prev, SsmSrc next, SsmSrc up, double* __restrict__ l1,
                                   double* __restrict__ l2) {
    const long total = a.T * a.Q;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const long t = idx / a.Q;
        const int q = (int)(idx % a.Q);
        double f1, f2, p1, p2, c1, c2, w;
        ssm_naturals(a, t, q, f1, f2, p1, p2, c1, c2, w);
        double a1 = 0.0, a2 = 0.0;
        const SsmSrc* srcs[3] = {&prev, &next, &up};
        GPB_UNROLL
        for (int s = 0; s < 3; s++) {
            const SsmSrc& S = *srcs[s];
            if (S.dm && t >= S.first && t < S.first + S.count) {
                const double dmc = S.sign * S.dm[(t - S.first) * S.ld + q], dvc = S.dv[(t - S.first) * S.ld + q];
                a1 += dmc / c2;
                a2 += -dmc * c1 / (c2 * c2) - dvc / (c2 * c2);
            }
        }
        l1[idx] = a1;
        l2[idx] = a2;
    }
}

// gradients wrt x_factor_1/2 and the posterior / cavity log-partition sums over ALL T rows
// (aep_models.py:1208-1232, 1287-1315, 1389-1437); part[block] = {phi_post, phi_cav}
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) ssm_xfinal_kernel(SsmArgs a, const double* __restrict__ l1,
                                                        const double* __restrict__ l2, double* __restrict__ gx1,
                                                        double* __restrict__ gx2, double* __restrict__ part) {
    GPB_SHARED double scratch[8];
    const long total = a.T * a.Q;
    double s0 = 0.0, s1 = 0.0;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const long t = idx / a.Q;
        const int q = (int)(idx % a.Q);
        double f1, f2, p1, p2, c1, c2, w3;
        ssm_naturals(a, t, q, f1, f2, p1, p2, c1, c2, w3);
        // per-row weights of the posterior / cavity terms: -(1 - 1/alpha) + 1/alpha per adjacent transition
        double sp = -(1.0 - 1.0 / a.alpha), sc = -1.0 / a.alpha;
        if (t < a.T - 1) { sp += 1.0 / a.alpha; sc += -1.0 / a.alpha; }
        if (t > 0) { sp += 1.0 / a.alpha; sc += -1.0 / a.alpha; }
        const double gp1 = sp * (p1 / p2);
        const double gp2 = sp * (-0.5 * p1 * p1 / (p2 * p2) - 0.5 / p2);
        double g1 = w3 * gp1;
        double g2 = 2.0 * w3 * gp2 * f2;
        const double w = w3 - a.alpha;
        g1 = g1 + (sc * (c1 / c2) + l1[idx]) * w;
        g2 = g2 + (sc * (-0.5 * c1 * c1 / (c2 * c2) - 0.5 / c2) + l2[idx]) * w * 2.0 * f2;
        gx1[idx] = g1;
        gx2[idx] = g2;
        s0 += sp * 0.5 * (p1 * p1 / p2 - log(p2));
        s1 += sc * 0.5 * (c1 * c1 / c2 - log(c2));
    }
    const double r0 = block_sum(s0, scratch);
    const double r1 = block_sum(s1, scratch);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = r0;
        part[2 * blockIdx.x + 1] = r1;
    }
}

}  // namespace gpb
