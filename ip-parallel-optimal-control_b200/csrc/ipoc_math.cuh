// Small fixed-size FP64 linear algebra and the two scan operators of the par IP-Newton path.
//
// Everything here is register-resident and fully unrolled for NX <= 8, NU <= NX; there are no
// tensor cores on this path (blocks are <= 8x8 FP64).  Formulas follow the conditional
// value-function elements (A, b, C, eta, J) of Sarkka & Garcia-Fernandez (IEEE TAC 2023) that
// the reference's `paroc.par_bwd_pass` scans (call site ref noc/par_interior_point_newton.py:120),
// and the affine-map composition of ref noc/costates.py:6-12.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define IPOC_DEV __device__ __forceinline__

namespace ipoc {

// ---------------------------------------------------------------- packed symmetric storage
template <int N>
struct Sym {
    static constexpr int SZ = N * (N + 1) / 2;
    IPOC_DEV static constexpr int at(int i, int j) {
        return (i <= j) ? (i * (2 * N - i + 1)) / 2 + (j - i) : (j * (2 * N - j + 1)) / 2 + (i - j);
    }
};

// ---------------------------------------------------------------- dense solve, partial pivoting
// In-place Gaussian elimination of W with row swaps applied to the right-hand sides X; the same
// pivot rule as LAPACK dgetf2 (first entry of maximum modulus) so that results track a
// host-side `solve` to rounding.  Row swaps are predicated selects on compile-time indices, so
// the arrays stay in registers.
template <int N, int R>
IPOC_DEV void lu_solve(double (&W)[N][N], double (&X)[N][R]) {
    double inv[N];   // reciprocal pivots: one FP64 reciprocal per column instead of a division per entry
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (k < N - 1) {
            int p = k;
            double best = fabs(W[k][k]);
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const double v = fabs(W[i][k]);
                if (v > best) { best = v; p = i; }
            }
            if (p != k) {
#pragma unroll
                for (int i = k + 1; i < N; ++i) {
                    const bool sw = (p == i);
#pragma unroll
                    for (int j = k; j < N; ++j) {
                        const double a = W[k][j], b = W[i][j];
                        W[k][j] = sw ? b : a;
                        W[i][j] = sw ? a : b;
                    }
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const double a = X[k][j], b = X[i][j];
                        X[k][j] = sw ? b : a;
                        X[i][j] = sw ? a : b;
                    }
                }
            }
        }
        inv[k] = 1.0 / W[k][k];
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const double m = W[i][k] * inv[k];
#pragma unroll
            for (int j = k + 1; j < N; ++j) W[i][j] -= m * W[k][j];
#pragma unroll
            for (int j = 0; j < R; ++j) X[i][j] -= m * X[k][j];
        }
    }
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
#pragma unroll
        for (int j = 0; j < R; ++j) {
            double s = X[k][j];
#pragma unroll
            for (int i = k + 1; i < N; ++i) s -= W[k][i] * X[i][j];
            X[k][j] = s * inv[k];
        }
    }
}

// Factor-once / substitute-many form of the same elimination (same pivot rule, same operation
// order per right-hand side, so results are bit-identical to lu_solve): W is overwritten by its LU
// factors (unit-lower multipliers below the diagonal), inv = reciprocal pivots, perm[i] = original
// index of the row now in position i.  The caller builds its right-hand sides directly in pivot
// order (row i of X = original row perm[i]) — a runtime-indexed load of the operands instead of a
// cascade of predicated register swaps over every right-hand-side column — and calls lu_subst.
template <int N>
IPOC_DEV void lu_factor(double (&W)[N][N], double (&inv)[N], int (&perm)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) perm[i] = i;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (k < N - 1) {
            int p = k;
            double best = fabs(W[k][k]);
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const double v = fabs(W[i][k]);
                if (v > best) { best = v; p = i; }
            }
            if (p != k) {
#pragma unroll
                for (int i = k + 1; i < N; ++i) {
                    const bool sw = (p == i);
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        const double a = W[k][j], b = W[i][j];
                        W[k][j] = sw ? b : a;
                        W[i][j] = sw ? a : b;
                    }
                    const int pa = perm[k], pb = perm[i];
                    perm[k] = sw ? pb : pa;
                    perm[i] = sw ? pa : pb;
                }
            }
        }
        inv[k] = 1.0 / W[k][k];
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const double m = W[i][k] * inv[k];
            W[i][k] = m;
#pragma unroll
            for (int j = k + 1; j < N; ++j) W[i][j] -= m * W[k][j];
        }
    }
}
template <int N, int R>
IPOC_DEV void lu_subst(const double (&W)[N][N], const double (&inv)[N], double (&X)[N][R]) {
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
        for (int i = k + 1; i < N; ++i)
#pragma unroll
            for (int j = 0; j < R; ++j) X[i][j] -= W[i][k] * X[k][j];
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
#pragma unroll
        for (int j = 0; j < R; ++j) {
            double s = X[k][j];
#pragma unroll
            for (int i = k + 1; i < N; ++i) s -= W[k][i] * X[i][j];
            X[k][j] = s * inv[k];
        }
    }
}

// Positive-definiteness of a small symmetric matrix (leading principal minors via elimination
// without pivoting).  Any NaN makes it false.  Stands in for `all(eigh(G) > 0)`
// (ref noc/seq_interior_point_newton.py:52-53).
template <int N>
IPOC_DEV bool is_pos_def(const double (&G)[N][N]) {
    double W[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) W[i][j] = 0.5 * (G[i][j] + G[j][i]);
    bool ok = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        ok = ok && (W[k][k] > 0.0);
        const double inv = 1.0 / W[k][k];
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const double m = W[i][k] * inv;
#pragma unroll
            for (int j = k + 1; j < N; ++j) W[i][j] -= m * W[k][j];
        }
    }
    return ok;
}

// ================================================================ affine operator (K1, K3)
// Element = map x -> F x + c.  compose(first, second) = second o first
// (ref noc/costates.py:6-12: Fik = Fjk Fij, cik = Fjk cij + cjk).
template <int NX>
struct AffElem {
    static constexpr int ESZ = NX * NX + NX;
    double r[ESZ];
    IPOC_DEV double& F(int i, int j) { return r[i * NX + j]; }
    IPOC_DEV double F(int i, int j) const { return r[i * NX + j]; }
    IPOC_DEV double& c(int i) { return r[NX * NX + i]; }
    IPOC_DEV double c(int i) const { return r[NX * NX + i]; }
};

template <int NX>
struct AffVal {
    static constexpr int VSZ = NX;
    double r[VSZ];
};

// Strided read-only views of elements kept in (shared) memory: component c lives at p[c * s].
// The scan kernels keep both operands of a combine in shared memory so that only the result and
// the temporaries occupy registers (two register-resident operands + result would exceed 255).
template <int NX>
struct AffView {
    const double* p;
    int s;
    IPOC_DEV double F(int i, int j) const { return p[(i * NX + j) * s]; }
    IPOC_DEV double c(int i) const { return p[(NX * NX + i) * s]; }
};

template <int NX>
struct AffOp {
    using Elem = AffElem<NX>;
    using Val = AffVal<NX>;
    static constexpr const char* tag_mid_up = "k_mid_up_aff";
    static constexpr const char* tag_mid_down = "k_mid_down_aff";
    static constexpr const char* tag_top = "k_top_aff";
    IPOC_DEV static void identity(Elem& e) {
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
            for (int j = 0; j < NX; ++j) e.F(i, j) = (i == j) ? 1.0 : 0.0;
            e.c(i) = 0.0;
        }
    }
    using View = AffView<NX>;
    // out = second o first
    template <class E2, class E1>
    IPOC_DEV static void compose_t(Elem& out, const E2& first, const E1& second) {
        Elem t;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < NX; ++k) s += second.F(i, k) * first.F(k, j);
                t.F(i, j) = s;
            }
            double s = second.c(i);
#pragma unroll
            for (int k = 0; k < NX; ++k) s += second.F(i, k) * first.c(k);
            t.c(i) = s;
        }
        out = t;
    }
    IPOC_DEV static void compose(Elem& out, const Elem& first, const Elem& second) {
        compose_t<Elem, Elem>(out, first, second);
    }
    template <class E>
    IPOC_DEV static void apply_t(Val& out, const E& e, const Val& x) {
        Val t;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double s = e.c(i);
#pragma unroll
            for (int k = 0; k < NX; ++k) s += e.F(i, k) * x.r[k];
            t.r[i] = s;
        }
        out = t;
    }
    IPOC_DEV static void apply(Val& out, const Elem& e, const Val& x) { apply_t<Elem>(out, e, x); }
    // memory-to-memory forms used by the generic scan kernels (strided operands, see RicOp)
    IPOC_DEV static void compose_mm(double* dst, int ds, const double* first, int fs, const double* second, int ss) {
        Elem o;
        compose_t(o, View{first, fs}, View{second, ss});
#pragma unroll
        for (int c = 0; c < Elem::ESZ; ++c) dst[c * ds] = o.r[c];
    }
    IPOC_DEV static void apply_mm(Val& v, const double* e, int es) { apply_t(v, View{e, es}, v); }
};

// ================================================================ Riccati operator (K2)
// Element (A, b, C, eta, J) of a time segment i -> j: conditional value function
//   V(x_i, x_j) dual form; C, J symmetric (packed upper).  A suffix segment k -> end carries the
//   value function V_k(x) = 1/2 x' J x - eta' x.
template <int NX>
struct RicElem {
    static constexpr int SY = Sym<NX>::SZ;
    static constexpr int OA = 0, OB = NX * NX, OC = OB + NX, OE = OC + SY, OJ = OE + NX;
    static constexpr int ESZ = OJ + SY;
    double r[ESZ];
    IPOC_DEV double& A(int i, int j) { return r[OA + i * NX + j]; }
    IPOC_DEV double A(int i, int j) const { return r[OA + i * NX + j]; }
    IPOC_DEV double& b(int i) { return r[OB + i]; }
    IPOC_DEV double b(int i) const { return r[OB + i]; }
    IPOC_DEV double& C(int i, int j) { return r[OC + Sym<NX>::at(i, j)]; }
    IPOC_DEV double C(int i, int j) const { return r[OC + Sym<NX>::at(i, j)]; }
    IPOC_DEV double& eta(int i) { return r[OE + i]; }
    IPOC_DEV double eta(int i) const { return r[OE + i]; }
    IPOC_DEV double& J(int i, int j) { return r[OJ + Sym<NX>::at(i, j)]; }
    IPOC_DEV double J(int i, int j) const { return r[OJ + Sym<NX>::at(i, j)]; }
};

template <int NX>
struct RicView {
    using E = RicElem<NX>;
    const double* p;
    int s;
    IPOC_DEV double A(int i, int j) const { return p[(E::OA + i * NX + j) * s]; }
    IPOC_DEV double b(int i) const { return p[(E::OB + i) * s]; }
    IPOC_DEV double C(int i, int j) const { return p[(E::OC + Sym<NX>::at(i, j)) * s]; }
    IPOC_DEV double eta(int i) const { return p[(E::OE + i) * s]; }
    IPOC_DEV double J(int i, int j) const { return p[(E::OJ + Sym<NX>::at(i, j)) * s]; }
};

// Value function V(x) = 1/2 x' S x - v' x
template <int NX>
struct RicVal {
    static constexpr int SY = Sym<NX>::SZ;
    static constexpr int VSZ = SY + NX;
    double r[VSZ];
    IPOC_DEV double& S(int i, int j) { return r[Sym<NX>::at(i, j)]; }
    IPOC_DEV double S(int i, int j) const { return r[Sym<NX>::at(i, j)]; }
    IPOC_DEV double& v(int i) { return r[SY + i]; }
    IPOC_DEV double v(int i) const { return r[SY + i]; }
};

template <int NX>
struct RicOp {
    using Elem = RicElem<NX>;
    using Val = RicVal<NX>;
    static constexpr const char* tag_mid_up = "k_mid_up_ric";
    static constexpr const char* tag_mid_down = "k_mid_down_ric";
    static constexpr const char* tag_top = "k_top_ric";

    IPOC_DEV static void identity(Elem& e) {
#pragma unroll
        for (int i = 0; i < Elem::ESZ; ++i) e.r[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NX; ++i) e.A(i, i) = 1.0;
    }

    // The scan runs backwards in time: `first` is the LATER segment (j -> k, already holding the
    // value information of the end of the horizon), `second` the EARLIER one (i -> j).
    // out = combine(e1 = second, e2 = first), with W = I + C1 J2:
    //   A = A2 W^-1 A1            b = A2 W^-1 (b1 + C1 eta2) + b2      C = A2 W^-1 C1 A2' + C2
    //   eta = A1' W^-T (eta2 - J2 b1) + eta1                            J = A1' W^-T J2 A1 + J1
    // Only ONE factorisation is needed: by the push-through identity W^-T J2 = J2 W^-1, so with
    // XA = W^-1 A1 and xb = W^-1 b1 (two of the right-hand sides of the first solve)
    //   J = A1' (J2 XA) + J1,      eta = XA' eta2 - A1' (J2 xb) + eta1.
    using View = RicView<NX>;
    template <class E2, class E1>
    IPOC_DEV static void compose_t(Elem& out, const E2& first, const E1& second) {
        const E1& e1 = second;
        const E2& e2 = first;
        Elem o;
        double W[NX][NX], inv[NX];
        int perm[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                double w = (i == j) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < NX; ++k) w += e1.C(i, k) * e2.J(k, j);
                W[i][j] = w;
            }
        }
        lu_factor<NX>(W, inv, perm);
        // right-hand sides [A1 (NX) | b1 | C1 eta2 | C1 A2' (NX)], built in pivot order: row i is the
        // original row perm[i] of e1 (a runtime-indexed read of the operand view)
        double X[NX][2 * NX + 2];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const int r = perm[i];
            double c1[NX];
#pragma unroll
            for (int k = 0; k < NX; ++k) c1[k] = e1.C(r, k);
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                double ca = 0.0;
#pragma unroll
                for (int k = 0; k < NX; ++k) ca += c1[k] * e2.A(j, k);   // (C1 A2')_{rj}
                X[i][j] = e1.A(r, j);
                X[i][NX + 2 + j] = ca;
            }
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < NX; ++k) s += c1[k] * e2.eta(k);
            X[i][NX] = e1.b(r);
            X[i][NX + 1] = s;
        }
        lu_subst<NX, 2 * NX + 2>(W, inv, X);
        double JX[NX][NX + 1];   // J2 [XA | xb]
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                double a = 0.0, jx = 0.0;
#pragma unroll
                for (int k = 0; k < NX; ++k) {
                    a += e2.A(i, k) * X[k][j];
                    jx += e2.J(i, k) * X[k][j];
                }
                o.A(i, j) = a;
                JX[i][j] = jx;
                if (j >= i) {
                    double c = e2.C(i, j);
#pragma unroll
                    for (int k = 0; k < NX; ++k) c += e2.A(i, k) * X[k][NX + 2 + j];
                    o.C(i, j) = c;
                }
            }
            double s = e2.b(i), jb = 0.0;
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                s += e2.A(i, k) * (X[k][NX] + X[k][NX + 1]);
                jb += e2.J(i, k) * X[k][NX];
            }
            o.b(i) = s;
            JX[i][NX] = jb;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double s = e1.eta(i);
#pragma unroll
            for (int k = 0; k < NX; ++k) s += X[k][i] * e2.eta(k) - e1.A(k, i) * JX[k][NX];
            o.eta(i) = s;
#pragma unroll
            for (int j = i; j < NX; ++j) {
                double c = e1.J(i, j);
#pragma unroll
                for (int k = 0; k < NX; ++k) c += e1.A(k, i) * JX[k][j];
                o.J(i, j) = c;
            }
        }
        out = o;
    }

    // Push a value function (S, v) at the end of segment e back to its start:
    //   S' = A' (I + S C)^-1 S A + J,   v' = A' (I + S C)^-1 (v - S b) + eta
    IPOC_DEV static void apply(Val& out, const Elem& e, const Val& in) { apply_t<Elem>(out, e, in); }
    // Memory-to-memory forms, deliberately NOT inlined: the scan kernels of the latency regime run
    // each combine only a handful of times with one or two warps per SM, where instruction fetch of
    // the ~2k-instruction unrolled body dominates (profiles/r01b: "no_instructions" stalls, 10
    // cycles per instruction); one shared copy per kernel keeps the fetched code small.
    // Operands are strided views (component c at p[c * stride]) in shared, local or global memory;
    // dst may alias an operand (the result is formed in registers first).
    static __device__ __noinline__ void compose_mm(double* dst, int ds, const double* first, int fs,
                                                   const double* second, int ss) {
        Elem o;
        compose_t(o, View{first, fs}, View{second, ss});
#pragma unroll
        for (int c = 0; c < Elem::ESZ; ++c) dst[c * ds] = o.r[c];
    }
    static __device__ __noinline__ void apply_mm(Val& v, const double* e, int es) { apply_t(v, View{e, es}, v); }
    template <class E>
    IPOC_DEV static void apply_t(Val& out, const E& e, const Val& in) {
        double W[NX][NX];
        double Y[NX][NX + 1];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                double w = (i == j) ? 1.0 : 0.0;
                double sa = 0.0;
#pragma unroll
                for (int k = 0; k < NX; ++k) {
                    w += in.S(i, k) * e.C(k, j);
                    sa += in.S(i, k) * e.A(k, j);
                }
                W[i][j] = w;
                Y[i][1 + j] = sa;
            }
            double s = in.v(i);
#pragma unroll
            for (int k = 0; k < NX; ++k) s -= in.S(i, k) * e.b(k);
            Y[i][0] = s;
        }
        lu_solve<NX, NX + 1>(W, Y);
        Val o;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double s = e.eta(i);
#pragma unroll
            for (int k = 0; k < NX; ++k) s += e.A(k, i) * Y[k][0];
            o.v(i) = s;
#pragma unroll
            for (int j = i; j < NX; ++j) {
                double c = e.J(i, j);
#pragma unroll
                for (int k = 0; k < NX; ++k) c += e.A(k, i) * Y[k][1 + j];
                o.S(i, j) = c;
            }
        }
        out = o;
    }
};


// ---------------------------------------------------------------- lane-cooperative Riccati combine
// The latency-critical level scans run ONE combine on NX cooperating lanes instead of one: lane g owns column g
// of every matrix result.  Both operands sit in shared memory (component c of a slot at p[c * 32]) and every
// lane reads what it needs from there (same-address reads of a group are broadcasts), so the lanes exchange
// nothing: the NX x NX factorisation of W = I + C1 J2 and the solve for b1 are done redundantly by every lane
// (cheap next to the 2 NX + 2 right-hand sides of the single-thread form), each lane then substitutes only ITS
// columns of [A1 | C1 A2'] — lane 0 also the two vector columns — and forms column g of A, C, J and entry g of
// eta.  About 1/3 of the instructions per lane of RicOp::compose_t, the same operations per output element.
template <int NX>
IPOC_DEV void ric_compose_coop(double* dst, const double* first, const double* second, int g) {
    using E = RicElem<NX>;
    const RicView<NX> e2{first, 32};    // the LATER segment (earlier in scan order)
    const RicView<NX> e1{second, 32};   // the EARLIER segment
    double W[NX][NX], inv[NX];
    int perm[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            double w = (i == j) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < NX; ++k) w += e1.C(i, k) * e2.J(k, j);
            W[i][j] = w;
        }
    }
    lu_factor<NX>(W, inv, perm);
    // right-hand sides in pivot order: [A1(:,g) | (C1 A2')(:,g) | b1 | C1 eta2]
    double X[NX][4];
    double a2row[NX];   // row g of A2
#pragma unroll
    for (int k = 0; k < NX; ++k) a2row[k] = e2.A(g, k);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        const int r = perm[i];
        double ca = 0.0, ce = 0.0;
#pragma unroll
        for (int k = 0; k < NX; ++k) {
            const double c = e1.C(r, k);
            ca += c * a2row[k];
            ce += c * e2.eta(k);
        }
        X[i][0] = e1.A(r, g);
        X[i][1] = ca;
        X[i][2] = e1.b(r);
        X[i][3] = ce;
    }
    lu_subst<NX, 4>(W, inv, X);
    // column g of A = A2 XA, of J2 XA, and J2 xb
    double jxa[NX], jxb[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        double a = 0.0, ja = 0.0, jb = 0.0;
#pragma unroll
        for (int k = 0; k < NX; ++k) {
            a += e2.A(i, k) * X[k][0];
            ja += e2.J(i, k) * X[k][0];
            jb += e2.J(i, k) * X[k][2];
        }
        dst[(E::OA + i * NX + g) * 32] = a;
        jxa[i] = ja;
        jxb[i] = jb;
    }
    // upper part of column g of C = C2 + A2 XD and J = J1 + A1' (J2 XA)
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        if (i <= g) {
            double c = e2.C(i, g), jj = e1.J(i, g);
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                c += e2.A(i, k) * X[k][1];
                jj += e1.A(k, i) * jxa[k];
            }
            dst[(E::OC + Sym<NX>::at(i, g)) * 32] = c;
            dst[(E::OJ + Sym<NX>::at(i, g)) * 32] = jj;
        }
    }
    // eta_g = eta1_g + XA(:,g)' eta2 - A1(:,g)' (J2 xb)
    double s = e1.eta(g);
#pragma unroll
    for (int k = 0; k < NX; ++k) s += X[k][0] * e2.eta(k) - e1.A(k, g) * jxb[k];
    dst[(E::OE + g) * 32] = s;
    if (g == 0) {   // b = b2 + A2 (xb + xc)
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double bb = e2.b(i);
#pragma unroll
            for (int k = 0; k < NX; ++k) bb += e2.A(i, k) * (X[k][2] + X[k][3]);
            dst[(E::OB + i) * 32] = bb;
        }
    }
}

// ================================================================ one time step of the LQ problem
// x+ = A x + B u + c,  stage cost 1/2 x'Xx + 1/2 u'Uu + x'Mu + q'x + p'u   (X, U symmetric)
template <int NX, int NU>
struct StepLQ {
    double A[NX][NX];
    double B[NX][NU];
    double c[NX];
    double X[Sym<NX>::SZ];
    double U[NU][NU];
    double M[NX][NU];
    double q[NX];
    double p[NU];
};

// Solve U Y = RHS for a small symmetric-ish U (general LU; a plain division when NU == 1).
template <int NU, int R>
IPOC_DEV void small_solve(const double (&U)[NU][NU], double (&Y)[NU][R]) {
    if constexpr (NU == 1) {
        const double inv = 1.0 / U[0][0];
#pragma unroll
        for (int j = 0; j < R; ++j) Y[0][j] = Y[0][j] * inv;
    } else {
        double W[NU][NU];
#pragma unroll
        for (int i = 0; i < NU; ++i)
#pragma unroll
            for (int j = 0; j < NU; ++j) W[i][j] = U[i][j];
        lu_solve<NU, R>(W, Y);
    }
}

// Single-step element in the form the backward fold consumes:
//   A1 = A - B U^-1 M',  b1 = c - B U^-1 p,  (C1 = B U^-1 B' kept implicit as B, U),
//   J1 = X - M U^-1 M',  eta1 = -q + M U^-1 p
template <int NX, int NU>
struct StepElem {
    double A1[NX][NX];
    double B[NX][NU];
    double U[NU][NU];
    double b1[NX];
    double J1[Sym<NX>::SZ];
    double eta1[NX];
};

template <int NX, int NU>
IPOC_DEV void make_step_elem(StepElem<NX, NU>& e, const StepLQ<NX, NU>& s) {
    // Y = U^-1 [M' | p]
    double Y[NU][NX + 1];
#pragma unroll
    for (int a = 0; a < NU; ++a) {
#pragma unroll
        for (int j = 0; j < NX; ++j) Y[a][j] = s.M[j][a];
        Y[a][NX] = s.p[a];
    }
    small_solve<NU, NX + 1>(s.U, Y);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            double a1 = s.A[i][j];
#pragma unroll
            for (int a = 0; a < NU; ++a) a1 -= s.B[i][a] * Y[a][j];
            e.A1[i][j] = a1;
            if (j >= i) {
                double jj = s.X[Sym<NX>::at(i, j)];
#pragma unroll
                for (int a = 0; a < NU; ++a) jj -= s.M[i][a] * Y[a][j];
                e.J1[Sym<NX>::at(i, j)] = jj;
            }
        }
        double bb = s.c[i], ee = -s.q[i];
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            bb -= s.B[i][a] * Y[a][NX];
            ee += s.M[i][a] * Y[a][NX];
            e.B[i][a] = s.B[i][a];
        }
        e.b1[i] = bb;
        e.eta1[i] = ee;
    }
#pragma unroll
    for (int a = 0; a < NU; ++a)
#pragma unroll
        for (int b = 0; b < NU; ++b) e.U[a][b] = s.U[a][b];
}

// Element of a single step as a full RicElem (start of a backward fold).
template <int NX, int NU>
IPOC_DEV void step_to_elem(RicElem<NX>& g, const StepElem<NX, NU>& e) {
    double Y[NU][NX];   // U^-1 B'
#pragma unroll
    for (int a = 0; a < NU; ++a)
#pragma unroll
        for (int j = 0; j < NX; ++j) Y[a][j] = e.B[j][a];
    small_solve<NU, NX>(e.U, Y);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            g.A(i, j) = e.A1[i][j];
            if (j >= i) {
                double c = 0.0;
#pragma unroll
                for (int a = 0; a < NU; ++a) c += e.B[i][a] * Y[a][j];
                g.C(i, j) = c;
                g.J(i, j) = e.J1[Sym<NX>::at(i, j)];
            }
        }
        g.b(i) = e.b1[i];
        g.eta(i) = e.eta1[i];
    }
}

// g <- combine(e1 = single step, e2 = g)   (prepend one EARLIER step to the aggregate g).
// C1 = B U^-1 B' has rank NU, so W^-1 = (I + C1 J2)^-1 = I - B Gam^-1 (J2 B)',  Gam = U + B' J2 B
// (Woodbury): no NX x NX factorisation is needed.
template <int NX, int NU>
IPOC_DEV void ric_prepend_step(RicElem<NX>& g, const StepElem<NX, NU>& e) {
    double Wm[NX][NU];   // J2 B
    double AB[NX][NU];   // A2 B
    double z[NX];        // eta2 - J2 b1
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            double w = 0.0, ab = 0.0;
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                w += g.J(i, k) * e.B[k][a];
                ab += g.A(i, k) * e.B[k][a];
            }
            Wm[i][a] = w;
            AB[i][a] = ab;
        }
        double s = g.eta(i);
#pragma unroll
        for (int k = 0; k < NX; ++k) s -= g.J(i, k) * e.b1[k];
        z[i] = s;
    }
    double Gam[NU][NU];
    // right-hand sides: [Wm' A1 | (A2 B)' | Wm' | B' eta2 - Wm' b1 | B' z]
    double Y[NU][3 * NX + 2];
#pragma unroll
    for (int a = 0; a < NU; ++a) {
#pragma unroll
        for (int b = 0; b < NU; ++b) {
            double s = e.U[a][b];
#pragma unroll
            for (int i = 0; i < NX; ++i) s += e.B[i][a] * Wm[i][b];
            Gam[a][b] = s;
        }
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < NX; ++k) s += Wm[k][a] * e.A1[k][j];
            Y[a][j] = s;
            Y[a][NX + j] = AB[j][a];
            Y[a][2 * NX + j] = Wm[j][a];
        }
        double s4 = 0.0, s5 = 0.0;
#pragma unroll
        for (int k = 0; k < NX; ++k) {
            s4 += e.B[k][a] * g.eta(k) - Wm[k][a] * e.b1[k];
            s5 += e.B[k][a] * z[k];
        }
        Y[a][3 * NX] = s4;
        Y[a][3 * NX + 1] = s5;
    }
    small_solve<NU, 3 * NX + 2>(Gam, Y);

    RicElem<NX> o;
    double WA[NX][NX];   // W^-1 A1
    double Jm[Sym<NX>::SZ];   // J2 - Wm Gam^-1 Wm'
    double zz[NX];       // W^-T z
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            double s = e.A1[i][j];
#pragma unroll
            for (int a = 0; a < NU; ++a) s -= e.B[i][a] * Y[a][j];
            WA[i][j] = s;
            if (j >= i) {
                double c = g.C(i, j), jm = g.J(i, j);
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    c += AB[i][a] * Y[a][NX + j];
                    jm -= Wm[i][a] * Y[a][2 * NX + j];
                }
                o.C(i, j) = c;
                Jm[Sym<NX>::at(i, j)] = jm;
            }
        }
        double bb = g.b(i), s = z[i];
#pragma unroll
        for (int k = 0; k < NX; ++k) bb += g.A(i, k) * e.b1[k];
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            bb += AB[i][a] * Y[a][3 * NX];
            s -= Wm[i][a] * Y[a][3 * NX + 1];
        }
        o.b(i) = bb;
        zz[i] = s;
    }
    double T1[NX][NX];   // Jm A1
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            double a = 0.0, t = 0.0;
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                a += g.A(i, k) * WA[k][j];
                t += Jm[Sym<NX>::at(i, k)] * e.A1[k][j];
            }
            o.A(i, j) = a;
            T1[i][j] = t;
        }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        double s = e.eta1[i];
#pragma unroll
        for (int k = 0; k < NX; ++k) s += e.A1[k][i] * zz[k];
        o.eta(i) = s;
#pragma unroll
        for (int j = i; j < NX; ++j) {
            double c = e.J1[Sym<NX>::at(i, j)];
#pragma unroll
            for (int k = 0; k < NX; ++k) c += e.A1[k][i] * T1[k][j];
            o.J(i, j) = c;
        }
    }
    g = o;
}

// One step of the seeded Riccati recursion (backward), producing the gains of that step:
//   G = U + B'S+B,  Kx = G^-1 (M' + B'S+A),  d = G^-1 (-p + B'(v+ - S+ c))     [u = -Kx x + d]
//   Fcl = A - B Kx,  ccl = c + B d
//   S = X + A' S+ Fcl - M Kx (symmetrised),  v = -q + Fcl'(v+ - S+ c) + Kx' p
// Returns d'Gd (for pred_reduction = -1/2 sum d'Gd) and whether G is positive definite.
template <int NX, int NU>
struct StepGain {
    double Kx[NU][NX];
    double d[NU];
    double Fcl[NX][NX];
    double ccl[NX];
    double dGd;
    bool pd;
};

template <int NX, int NU>
IPOC_DEV void ric_step_back(RicVal<NX>& val, StepGain<NX, NU>& o, const StepLQ<NX, NU>& s) {
    double Wm[NX][NU];   // S+ B
    double w[NX];        // v+ - S+ c
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int a = 0; a < NU; ++a) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < NX; ++k) t += val.S(i, k) * s.B[k][a];
            Wm[i][a] = t;
        }
        double t = val.v(i);
#pragma unroll
        for (int k = 0; k < NX; ++k) t -= val.S(i, k) * s.c[k];
        w[i] = t;
    }
    double G[NU][NU];
    double Y[NU][NX + 1];
#pragma unroll
    for (int a = 0; a < NU; ++a) {
#pragma unroll
        for (int b = 0; b < NU; ++b) {
            double t = s.U[a][b];
#pragma unroll
            for (int i = 0; i < NX; ++i) t += s.B[i][a] * Wm[i][b];
            G[a][b] = t;
        }
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            double t = s.M[j][a];
#pragma unroll
            for (int k = 0; k < NX; ++k) t += Wm[k][a] * s.A[k][j];
            Y[a][j] = t;
        }
        double t = -s.p[a];
#pragma unroll
        for (int k = 0; k < NX; ++k) t += s.B[k][a] * w[k];
        Y[a][NX] = t;
    }
    o.pd = is_pos_def<NU>(G);
    small_solve<NU, NX + 1>(G, Y);
    double dGd = 0.0;
#pragma unroll
    for (int a = 0; a < NU; ++a) {
#pragma unroll
        for (int j = 0; j < NX; ++j) o.Kx[a][j] = Y[a][j];
        o.d[a] = Y[a][NX];
    }
#pragma unroll
    for (int a = 0; a < NU; ++a) {
        double t = 0.0;
#pragma unroll
        for (int b = 0; b < NU; ++b) t += G[a][b] * o.d[b];
        dGd += o.d[a] * t;
    }
    o.dGd = dGd;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            double t = s.A[i][j];
#pragma unroll
            for (int a = 0; a < NU; ++a) t -= s.B[i][a] * o.Kx[a][j];
            o.Fcl[i][j] = t;
        }
        double t = s.c[i];
#pragma unroll
        for (int a = 0; a < NU; ++a) t += s.B[i][a] * o.d[a];
        o.ccl[i] = t;
    }
    double SF[NX][NX];   // S+ Fcl
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < NX; ++k) t += val.S(i, k) * o.Fcl[k][j];
            SF[i][j] = t;
        }
    RicVal<NX> nv;
    double full[NX][NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < NX; ++k) t += s.A[k][i] * SF[k][j];
#pragma unroll
            for (int a = 0; a < NU; ++a) t -= s.M[i][a] * o.Kx[a][j];
            full[i][j] = t;
        }
        double t = -s.q[i];
#pragma unroll
        for (int k = 0; k < NX; ++k) t += o.Fcl[k][i] * w[k];
#pragma unroll
        for (int a = 0; a < NU; ++a) t += o.Kx[a][i] * s.p[a];
        nv.v(i) = t;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
        for (int j = i; j < NX; ++j) nv.S(i, j) = s.X[Sym<NX>::at(i, j)] + 0.5 * (full[i][j] + full[j][i]);
    val = nv;
}

}  // namespace ipoc
