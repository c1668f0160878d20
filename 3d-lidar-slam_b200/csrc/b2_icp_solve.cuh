// b2_icp_solve.cuh — the per-step SE(3) solve: Umeyama rotation (Newton polar iteration, Jacobi SVD fallback),
// the point-to-plane 6x6 solve and the A.3 bookkeeping (fitness, rmse, convergence, T <- dT * T).
// Included by b2_icp.cu inside namespace b2 { namespace {.
#pragma once
// ---------------------------------------------------------------------------
// small dense solvers (single thread, fp64)
// ---------------------------------------------------------------------------

// Rotation maximising tr(R * C^T) for C = sum (q - mq)(p - mp)^T / n: R = U diag(1,1,det(U)det(V)) V^T.
// One-sided (Hestenes) Jacobi: columns of A = C*V are driven orthogonal; U's third column is never
// taken from data (it is u1 x u2), which keeps R a rotation when the correspondences are coplanar.
__device__ __noinline__ void rotation_from_covariance(const double C[9], double R[9]) {
    double A[9], V[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        A[i] = C[i];
        V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0;
            const int q = pq == 0 ? 1 : 2;
            double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                alpha += A[3 * k + p] * A[3 * k + p];
                beta += A[3 * k + q] * A[3 * k + q];
                gamma += A[3 * k + p] * A[3 * k + q];
            }
            // 8 ulp: below that the columns are orthogonal to working precision (a tighter test
            // never settles — rounding noise keeps it rotating for all 60 sweeps)
            if (gamma != 0.0 && gamma * gamma > 3e-30 * (alpha * beta)) {
                rotated = true;
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double ap = A[3 * k + p], aq = A[3 * k + q];
                    A[3 * k + p] = cs * ap - sn * aq;
                    A[3 * k + q] = sn * ap + cs * aq;
                    double vp = V[3 * k + p], vq = V[3 * k + q];
                    V[3 * k + p] = cs * vp - sn * vq;
                    V[3 * k + q] = sn * vp + cs * vq;
                }
            }
        }
        if (!rotated) break;
    }
    double s[3];
    int ord[3] = {0, 1, 2};
#pragma unroll
    for (int i = 0; i < 3; ++i) s[i] = sqrt(A[i] * A[i] + A[3 + i] * A[3 + i] + A[6 + i] * A[6 + i]);
    if (s[ord[0]] < s[ord[1]]) { int t = ord[0]; ord[0] = ord[1]; ord[1] = t; }
    if (s[ord[1]] < s[ord[2]]) { int t = ord[1]; ord[1] = ord[2]; ord[2] = t; }
    if (s[ord[0]] < s[ord[1]]) { int t = ord[0]; ord[0] = ord[1]; ord[1] = t; }
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    const double s0 = s[ord[0]], s1 = s[ord[1]];
    if (!(s0 > 0.0)) return;
    double u1[3], u2[3], u3[3], v1[3], v2[3], v3[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        u1[k] = A[3 * k + ord[0]] / s0;
        v1[k] = V[3 * k + ord[0]];
        v2[k] = V[3 * k + ord[1]];
        v3[k] = V[3 * k + ord[2]];
    }
    if (s1 > 1e-14 * s0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) u2[k] = A[3 * k + ord[1]] / s1;
    } else {  // rank one: any unit vector orthogonal to u1
        int m = fabs(u1[0]) <= fabs(u1[1]) ? (fabs(u1[0]) <= fabs(u1[2]) ? 0 : 2) : (fabs(u1[1]) <= fabs(u1[2]) ? 1 : 2);
        u2[0] = u2[1] = u2[2] = 0.0;
        u2[m] = 1.0;
    }
    double d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
    double nn = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        u2[k] -= d * u1[k];
        nn += u2[k] * u2[k];
    }
    nn = 1.0 / sqrt(nn);
#pragma unroll
    for (int k = 0; k < 3; ++k) u2[k] *= nn;
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1];
    u3[1] = u1[2] * u2[0] - u1[0] * u2[2];
    u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    const double detV = v1[0] * (v2[1] * v3[2] - v2[2] * v3[1]) - v2[0] * (v1[1] * v3[2] - v1[2] * v3[1]) +
                        v3[0] * (v1[1] * v2[2] - v1[2] * v2[1]);
    const double sg = detV < 0 ? -1.0 : 1.0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cidx = 0; cidx < 3; ++cidx)
            R[3 * r + cidx] = u1[r] * v1[cidx] + u2[r] * v2[cidx] + sg * u3[r] * v3[cidx];
}

// Gaussian elimination with partial pivoting on the symmetric 6x6 normal equations.
__device__ __noinline__ bool solve6(const double* JtJ_upper /*21*/, const double* rhs /*6*/, double x[6]) {
    double M[6][7];
    int t = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) {
            M[i][j] = JtJ_upper[t];
            M[j][i] = JtJ_upper[t];
            ++t;
        }
    for (int i = 0; i < 6; ++i) M[i][6] = rhs[i];
    for (int col = 0; col < 6; ++col) {
        int piv = col;
        for (int r = col + 1; r < 6; ++r)
            if (fabs(M[r][col]) > fabs(M[piv][col])) piv = r;
        if (!(fabs(M[piv][col]) > 0.0) || !isfinite(M[piv][col])) return false;
        if (piv != col)
            for (int k = 0; k < 7; ++k) {
                double tmp = M[col][k];
                M[col][k] = M[piv][k];
                M[piv][k] = tmp;
            }
        for (int r = col + 1; r < 6; ++r) {
            double f = M[r][col] / M[col][col];
            for (int k = col; k < 7; ++k) M[r][k] -= f * M[col][k];
        }
    }
    for (int i = 5; i >= 0; --i) {
        double v = M[i][6];
        for (int k = i + 1; k < 6; ++k) v -= M[i][k] * x[k];
        x[i] = v / M[i][i];
        if (!isfinite(x[i])) return false;
    }
    return true;
}

// Polar factor of a 3x3 with positive determinant by Frobenius-scaled Newton iteration,
// X <- (g X + (g X)^-T) / 2: when det(C) > 0 the orthogonal polar factor IS U V^T, the Umeyama
// rotation.  Everything stays in registers; the scale factor only steers convergence, so it is
// evaluated in float32.  Returns false (caller falls back to the Jacobi SVD) for det <= 0 — the
// reflection / rank-deficient cases — or if the iteration does not settle.
__device__ __forceinline__ bool polar_rotation(double a, double b, double c, double d, double e, double f, double g,
                                               double h, double i, double* R) {
    {
        const double nrm = sqrt(a * a + b * b + c * c + d * d + e * e + f * f + g * g + h * h + i * i);
        if (!(nrm > 0.0) || !isfinite(nrm)) return false;
        const double s = 1.7320508075688772 / nrm;
        a *= s; b *= s; c *= s; d *= s; e *= s; f *= s; g *= s; h *= s; i *= s;
    }
    bool ok = false;
#pragma unroll 1
    for (int it = 0; it < 40; ++it) {
        const double c00 = e * i - f * h, c01 = f * g - d * i, c02 = d * h - e * g;
        const double c10 = c * h - b * i, c11 = a * i - c * g, c12 = b * g - a * h;
        const double c20 = b * f - c * e, c21 = c * d - a * f, c22 = a * e - b * d;
        const double det = a * c00 + b * c01 + c * c02;
        if (!(det > 1e-12)) return false;  // (scaled so that |X|_F = sqrt(3): an orthogonal X has det 1)
        const double nx2 = a * a + b * b + c * c + d * d + e * e + f * f + g * g + h * h + i * i;
        const double nc2 = c00 * c00 + c01 * c01 + c02 * c02 + c10 * c10 + c11 * c11 + c12 * c12 + c20 * c20 +
                           c21 * c21 + c22 * c22;
        // gamma = (|X^-T|_F / |X|_F)^(1/2), |X^-T|_F = sqrt(nc2)/det
        const float gam = sqrtf(sqrtf((float)(nc2 / (det * det * nx2))));
        const double ga = 0.5 * (double)gam, gb = 0.5 / ((double)gam * det);
        const double na = ga * a + gb * c00, nb = ga * b + gb * c01, nc = ga * c + gb * c02;
        const double nd = ga * d + gb * c10, ne = ga * e + gb * c11, nf = ga * f + gb * c12;
        const double ng = ga * g + gb * c20, nh = ga * h + gb * c21, ni = ga * i + gb * c22;
        const double diff = (na - a) * (na - a) + (nb - b) * (nb - b) + (nc - c) * (nc - c) + (nd - d) * (nd - d) +
                            (ne - e) * (ne - e) + (nf - f) * (nf - f) + (ng - g) * (ng - g) + (nh - h) * (nh - h) +
                            (ni - i) * (ni - i);
        a = na; b = nb; c = nc; d = nd; e = ne; f = nf; g = ng; h = nh; i = ni;
        if (diff < 1e-22) {  // quadratic convergence: the next update is below 1e-20
            ok = true;
            break;
        }
    }
    if (!ok) return false;
    {  // one unscaled polishing step
        const double c00 = e * i - f * h, c01 = f * g - d * i, c02 = d * h - e * g;
        const double c10 = c * h - b * i, c11 = a * i - c * g, c12 = b * g - a * h;
        const double c20 = b * f - c * e, c21 = c * d - a * f, c22 = a * e - b * d;
        const double hd = 0.5 / (a * c00 + b * c01 + c * c02);
        R[0] = 0.5 * a + hd * c00; R[1] = 0.5 * b + hd * c01; R[2] = 0.5 * c + hd * c02;
        R[3] = 0.5 * d + hd * c10; R[4] = 0.5 * e + hd * c11; R[5] = 0.5 * f + hd * c12;
        R[6] = 0.5 * g + hd * c20; R[7] = 0.5 * h + hd * c21; R[8] = 0.5 * i + hd * c22;
    }
    return true;
}

// Umeyama rotation by Newton's method on SO(3), started at the identity.  The rotation maximises
// f(R) = tr(R^T C); with M = R^T C, S = sym(M) and a = axial(M - M^T), the second-order model of
// f(R exp[w]) is tr(M) + a.w - w^T (tr(S) I - S) w / 2, so the Newton step solves H w = a, H = tr(S) I - S, and
// R is retracted along the Cayley map (an exact rotation, equal to exp[w] to second order — the quadratic
// convergence is kept; one division per step serves both the 3x3 solve and the map).  An ICP update is a small
// rotation (1e-2 rad early, 1e-5 late), so two to four steps reach 1e-16 where the scaled polar iteration needs
// six to eight from a general matrix.  H positive definite at the fixed point is exactly the condition for
// the GLOBAL maximum over SO(3) (its eigenvalues are the pairwise sums of the signed singular values), i.e.
// for U diag(1,1,det(UV^T)) V^T, reflections included; anything else — H not positive definite, no
// convergence — returns false and the caller takes the polar / Jacobi path.
__device__ __forceinline__ bool procrustes_newton(const double C[9], double R[9]) {
    double r0 = 1, r1 = 0, r2 = 0, r3 = 0, r4 = 1, r5 = 0, r6 = 0, r7 = 0, r8 = 1;
    double m0 = C[0], m1 = C[1], m2 = C[2], m3 = C[3], m4 = C[4], m5 = C[5], m6 = C[6], m7 = C[7], m8 = C[8];
    const double scale2 = m0 * m0 + m1 * m1 + m2 * m2 + m3 * m3 + m4 * m4 + m5 * m5 + m6 * m6 + m7 * m7 + m8 * m8;
    if (!(scale2 > 0.0) || !isfinite(scale2)) return false;
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
        const double tr = m0 + m4 + m8;
        // H = tr(S) I - S (symmetric), S = (M + M^T) / 2
        const double h00 = tr - m0, h11 = tr - m4, h22 = tr - m8;
        const double h01 = -0.5 * (m1 + m3), h02 = -0.5 * (m2 + m6), h12 = -0.5 * (m5 + m7);
        const double a0 = m7 - m5, a1 = m2 - m6, a2 = m3 - m1;
        // adjugate of H and Sylvester's criterion
        const double c00 = h11 * h22 - h12 * h12, c01 = h02 * h12 - h01 * h22, c02 = h01 * h12 - h02 * h11;
        const double c11 = h00 * h22 - h02 * h02, c12 = h01 * h02 - h00 * h12, c22 = h00 * h11 - h01 * h01;
        const double det = h00 * c00 + h01 * c01 + h02 * c02;
        // (relative thresholds: a rank-deficient covariance must go to the SVD path, not through a tiny pivot)
        if (!(h00 > 1e-9 * tr) || !(c22 > 1e-9 * tr * tr) || !(det > 1e-9 * tr * tr * tr)) return false;
        // v = adj(H) a = det * w ;  u = w / 2 = v / (2 det) ;  Q = I + (4 det [v] + 2 [v]^2) / (4 det^2 + |v|^2)
        const double v0 = c00 * a0 + c01 * a1 + c02 * a2;
        const double v1 = c01 * a0 + c11 * a1 + c12 * a2;
        const double v2 = c02 * a0 + c12 * a1 + c22 * a2;
        const double vv = v0 * v0 + v1 * v1 + v2 * v2;
        const double d4 = 4.0 * det * det;
        const double inv = 1.0 / (d4 + vv);
        const double ka = 4.0 * det * inv, kb = 2.0 * inv;
        // [v]^2 = v v^T - |v|^2 I
        const double q0 = 1.0 + kb * (v0 * v0 - vv), q4 = 1.0 + kb * (v1 * v1 - vv), q8 = 1.0 + kb * (v2 * v2 - vv);
        const double q1 = kb * v0 * v1 - ka * v2, q3 = kb * v0 * v1 + ka * v2;
        const double q2 = kb * v0 * v2 + ka * v1, q6 = kb * v0 * v2 - ka * v1;
        const double q5 = kb * v1 * v2 - ka * v0, q7 = kb * v1 * v2 + ka * v0;
        // R <- R Q
        const double n0 = r0 * q0 + r1 * q3 + r2 * q6, n1 = r0 * q1 + r1 * q4 + r2 * q7, n2 = r0 * q2 + r1 * q5 + r2 * q8;
        const double n3 = r3 * q0 + r4 * q3 + r5 * q6, n4 = r3 * q1 + r4 * q4 + r5 * q7, n5 = r3 * q2 + r4 * q5 + r5 * q8;
        const double n6 = r6 * q0 + r7 * q3 + r8 * q6, n7 = r6 * q1 + r7 * q4 + r8 * q7, n8 = r6 * q2 + r7 * q5 + r8 * q8;
        r0 = n0; r1 = n1; r2 = n2; r3 = n3; r4 = n4; r5 = n5; r6 = n6; r7 = n7; r8 = n8;
        // |w|^2 = |v|^2 / det^2: below 1e-16 the step just applied leaves an error of ~|w|^2 (quadratic)
        if (vv < 1e-16 * det * det) {
            R[0] = r0; R[1] = r1; R[2] = r2; R[3] = r3; R[4] = r4; R[5] = r5; R[6] = r6; R[7] = r7; R[8] = r8;
            return true;
        }
        // M <- R^T C
        m0 = r0 * C[0] + r3 * C[3] + r6 * C[6]; m1 = r0 * C[1] + r3 * C[4] + r6 * C[7]; m2 = r0 * C[2] + r3 * C[5] + r6 * C[8];
        m3 = r1 * C[0] + r4 * C[3] + r7 * C[6]; m4 = r1 * C[1] + r4 * C[4] + r7 * C[7]; m5 = r1 * C[2] + r4 * C[5] + r7 * C[8];
        m6 = r2 * C[0] + r5 * C[3] + r8 * C[6]; m7 = r2 * C[1] + r5 * C[4] + r8 * C[7]; m8 = r2 * C[2] + r5 * C[5] + r8 * C[8];
    }
    return false;
}

// Evaluate one A.3 loop step from the reduced sums.  Runs on one thread; written so that every
// small matrix stays in registers (a single GPU thread retires roughly one dependent instruction
// per 6-10 cycles, so the instruction count of this tail is what the iteration latency is made of).
template <int MODE>
__device__ __noinline__ void icp_step(const double* S, int ns, const IcpParams prm, const IcpState& cur,
                                      IcpState* st, IcpResultDev* res, int* ndone) {
    const double n = MODE == 0 ? S[16] : S[28];
    const double sd2 = MODE == 0 ? S[15] : S[27];
    const double fitness = ns > 0 ? n / (double)ns : 0.0;
    const double rmse = n > 0 ? sqrt(sd2 / n) : 0.0;
    bool done = false;
    if (cur.has_prev && fabs(cur.prev_fitness - fitness) < prm.rel_fitness &&
        fabs(cur.prev_rmse - rmse) < prm.rel_rmse)
        done = true;
    if (cur.iter >= prm.max_iter) done = true;
    st->fitness = fitness;
    st->rmse = rmse;
    st->n_corr = (int)n;
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = cur.T[k];
    if (done) {
        st->done = 1;
#pragma unroll
        for (int k = 0; k < 12; ++k) res->T[k] = T[k];
        res->T[12] = 0.0; res->T[13] = 0.0; res->T[14] = 0.0; res->T[15] = 1.0;
        res->fitness = fitness;
        res->rmse = rmse;
        res->iterations = cur.iter;
        res->n_corr = (int)n;
        atomicAdd(ndone, 1);
        return;
    }
    double U[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};  // update, 3x4
    if (n > 0) {
        if (MODE == 0) {
            const double inv = 1.0 / n;
            double mp0 = S[0] * inv, mp1 = S[1] * inv, mp2 = S[2] * inv;
            double mq0 = S[3] * inv, mq1 = S[4] * inv, mq2 = S[5] * inv;
            double C[9], R[9];
            C[0] = S[6] * inv - mq0 * mp0;  C[1] = S[7] * inv - mq0 * mp1;  C[2] = S[8] * inv - mq0 * mp2;
            C[3] = S[9] * inv - mq1 * mp0;  C[4] = S[10] * inv - mq1 * mp1; C[5] = S[11] * inv - mq1 * mp2;
            C[6] = S[12] * inv - mq2 * mp0; C[7] = S[13] * inv - mq2 * mp1; C[8] = S[14] * inv - mq2 * mp2;
            if (!procrustes_newton(C, R) &&
                !polar_rotation(C[0], C[1], C[2], C[3], C[4], C[5], C[6], C[7], C[8], R))
                rotation_from_covariance(C, R);
            const double p0 = cur.pivot[0], p1 = cur.pivot[1], p2 = cur.pivot[2];
            mp0 += p0; mp1 += p1; mp2 += p2;
            mq0 += p0; mq1 += p1; mq2 += p2;
            U[0] = R[0]; U[1] = R[1]; U[2] = R[2];  U[3] = mq0 - (R[0] * mp0 + R[1] * mp1 + R[2] * mp2);
            U[4] = R[3]; U[5] = R[4]; U[6] = R[5];  U[7] = mq1 - (R[3] * mp0 + R[4] * mp1 + R[5] * mp2);
            U[8] = R[6]; U[9] = R[7]; U[10] = R[8]; U[11] = mq2 - (R[6] * mp0 + R[7] * mp1 + R[8] * mp2);
            st->pivot[0] = mq0; st->pivot[1] = mq1; st->pivot[2] = mq2;
        } else {
            double rhs[6], x[6];
            for (int k = 0; k < 6; ++k) rhs[k] = -S[21 + k];
            if (solve6(S, rhs, x)) {
                const double ca = cos(x[0]), sa = sin(x[0]), cb = cos(x[1]), sb = sin(x[1]);
                const double cg = cos(x[2]), sg = sin(x[2]);
                U[0] = cg * cb; U[1] = cg * sb * sa - sg * ca; U[2] = cg * sb * ca + sg * sa; U[3] = x[3];
                U[4] = sg * cb; U[5] = sg * sb * sa + cg * ca; U[6] = sg * sb * ca - cg * sa; U[7] = x[4];
                U[8] = -sb;     U[9] = cb * sa;                U[10] = cb * ca;               U[11] = x[5];
            }
        }
    }
    // T <- U * T for affine 3x4 blocks (the last row stays 0,0,0,1)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double u0 = U[4 * r], u1 = U[4 * r + 1], u2 = U[4 * r + 2], u3 = U[4 * r + 3];
        st->T[4 * r + 0] = u0 * T[0] + u1 * T[4] + u2 * T[8];
        st->T[4 * r + 1] = u0 * T[1] + u1 * T[5] + u2 * T[9];
        st->T[4 * r + 2] = u0 * T[2] + u1 * T[6] + u2 * T[10];
        st->T[4 * r + 3] = u0 * T[3] + u1 * T[7] + u2 * T[11] + u3;
    }
    st->iter = cur.iter + 1;
    st->prev_fitness = fitness;
    st->prev_rmse = rmse;
    st->has_prev = 1;
}
