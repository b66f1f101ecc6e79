// fans_cpu.cpp — multithreaded C++ restatement of the reference's LINEAR solve path, the CPU baseline of bench.py.
//
// TEST / MEASUREMENT INFRASTRUCTURE ONLY (like everything under oracle/): it is the checker and the timed CPU arm, never the product.
// The reference itself (MPI + FFTW + HDF5 + Eigen) cannot be built in this image, so this file restates, with std::thread over
// x-slabs instead of MPI ranks and its own radix-2 FFT instead of FFTW, exactly the loop the reference times at solver.h:292-298:
//   Solver ctor + computeFundamentalSolution      include/solver.h:105-204     -> Cpu::build_gamma
//   compute_residual_basic / iterateCubes         include/solver.h:229-385     -> Cpu::apply_elements
//   element_residual for a LinearModel            include/matmodel.h:190-201   -> K_phase ue + B^T C g0 v_e/n_gp (same numbers, K and the
//                                                                                  load vector are pre-multiplied like phase_stiffness)
//   convolution                                    include/solver.h:387-412     -> Cpu::convolution (r2c, Gamma multiply, c2r; 1/N in Gamma)
//   SolverCG::internalSolve, linear branch         include/solverCG.h:61-117    -> fcpu_solve_cg
//   compute_error                                  include/solver.h:414-452     -> norms
//   get_homogenized_stress                         include/solver.h:707-737     -> Cpu::homogenized_stress
// parity: pinned against oracle/fans_oracle.py (itself pinned by the reference's KATs) in tests/test_cpu_port.py.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

typedef std::complex<double> cplx;

namespace {

struct Pool {
    int nt;
    explicit Pool(int n) : nt(n < 1 ? 1 : n) {}
    // static block partition of [0, n) over the threads; fn(begin, end, thread)
    void run(long n, const std::function<void(long, long, int)> &fn) const
    {
        const int t = (int)std::min<long>(nt, n > 0 ? n : 1);
        if (t <= 1) {
            fn(0, n, 0);
            return;
        }
        std::vector<std::thread> th;
        th.reserve(t);
        for (int i = 0; i < t; ++i) th.emplace_back([&, i]() { fn(n * i / t, n * (i + 1) / t, i); });
        for (auto &x : th) x.join();
    }
};

// in-place radix-2 FFT of VL interleaved lines: a[k * VL + v], k < n (n a power of two); sign -1 forward, +1 inverse (unscaled)
template <int VL>
void fft_lines(cplx *a, int n, const cplx *tw /* exp(-2 pi i k / n), k < n/2 */, const int *rev, bool inverse)
{
    for (int i = 0; i < n; ++i) {
        const int j = rev[i];
        if (i < j)
            for (int v = 0; v < VL; ++v) std::swap(a[(size_t)i * VL + v], a[(size_t)j * VL + v]);
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < half; ++k) {
                cplx w = tw[k * step];
                if (inverse) w = std::conj(w);
                cplx *p = a + (size_t)(i + k) * VL, *q = a + (size_t)(i + k + half) * VL;
                for (int v = 0; v < VL; ++v) {
                    const double xr = q[v].real() * w.real() - q[v].imag() * w.imag();
                    const double xi = q[v].real() * w.imag() + q[v].imag() * w.real();
                    q[v] = cplx(p[v].real() - xr, p[v].imag() - xi);
                    p[v] = cplx(p[v].real() + xr, p[v].imag() + xi);
                }
            }
    }
}

struct Plan {
    int n = 0;
    std::vector<cplx> tw;
    std::vector<int> rev;
    void init(int n_)
    {
        n = n_;
        tw.resize(std::max(1, n / 2));
        for (int k = 0; k < n / 2; ++k) tw[k] = std::polar(1.0, -2.0 * M_PI * k / n);
        rev.assign(n, 0);
        int bits = 0;
        while ((1 << bits) < n) ++bits;
        for (int i = 0; i < n; ++i) {
            int r = 0;
            for (int b = 0; b < bits; ++b)
                if (i & (1 << b)) r |= 1 << (bits - 1 - b);
            rev[i] = r;
        }
    }
};

// symmetric h x h (h <= 3) pseudo-inverse with an ABSOLUTE cut on the eigenvalues (JacobiSVD pinv, solver.h:89-96, tol 1e-14)
void pinv_sym(int h, const double *A, double *out)
{
    double a[3][3] = {{0}}, v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < h; ++j) a[i][j] = A[i * h + j];
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < h; ++i)
            for (int j = i + 1; j < h; ++j) off += a[i][j] * a[i][j];
        if (off < 1e-300) break;
        for (int p = 0; p < h; ++p)
            for (int q = p + 1; q < h; ++q) {
                if (std::fabs(a[p][q]) < 1e-300) continue;
                const double th = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < h; ++k) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < h; ++k) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < h; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < h; ++j) {
            double s = 0.0;
            for (int k = 0; k < h; ++k)
                if (std::fabs(a[k][k]) > 1e-14) s += v[i][k] * v[j][k] / a[k][k];
            out[i * h + j] = s;
        }
}

struct Cpu {
    int nx, ny, nz, h, nstr, kzc, nd;
    double L[3], le[3], ve;
    size_t N;
    Pool pool;
    std::vector<uint16_t> ms;
    int nph = 0;
    std::vector<double> Kph, fgph, Cph;   // per phase: K (nd x nd), B^T C-part for the load vector (nd x nstr), C (nstr x nstr)
    std::vector<double> Bgp;              // [8][nstr][nd]
    std::vector<double> gamma;            // [x][y][kz][h*(h+1)/2]
    std::vector<cplx> spec;               // [x][y][kzc][h]
    std::vector<double> u, r, s, d, rnew;
    Plan px, py, pzh;                     // pzh: half-length plan of the real transform
    std::vector<cplx> twz;                // exp(-2 pi i k / nz), k <= nz/2
    double fft_seconds = 0.0;

    Cpu(int nx_, int ny_, int nz_, const double *L_, int h_, int nt) : nx(nx_), ny(ny_), nz(nz_), h(h_), pool(nt)
    {
        nstr = (h == 1) ? 3 : 6;
        nd = 8 * h;
        kzc = nz / 2 + 1;
        N = (size_t)nx * ny * nz;
        for (int i = 0; i < 3; ++i) L[i] = L_[i];
        le[0] = L[0] / nx, le[1] = L[1] / ny, le[2] = L[2] / nz;
        ve = le[0] * le[1] * le[2];
        px.init(nx), py.init(ny), pzh.init(nz / 2);
        twz.resize(kzc);
        for (int k = 0; k < kzc; ++k) twz[k] = std::polar(1.0, -2.0 * M_PI * k / nz);
        build_B();
        spec.resize((size_t)nx * ny * kzc * h);
        for (auto *f : {&u, &r, &s, &d, &rnew}) f->assign(N * h, 0.0);
    }

    // B at the 8 Gauss points (matmodel.h:104-188, 284-304), HEX8
    void build_B()
    {
        const double xp = 0.5 + std::sqrt(3.0) / 6.0, xm = 0.5 - std::sqrt(3.0) / 6.0, rs = 7.071067811865476e-01;
        Bgp.assign((size_t)8 * nstr * nd, 0.0);
        for (int g = 0; g < 8; ++g) {
            const double x = (g & 1) ? xp : xm, y = (g & 2) ? xp : xm, z = (g & 4) ? xp : xm;
            const double v0[8] = {-(1 - y) * (1 - z), (1 - y) * (1 - z), -y * (1 - z), y * (1 - z), -(1 - y) * z, (1 - y) * z, -y * z, y * z};
            const double v1[8] = {-(1 - x) * (1 - z), -x * (1 - z), (1 - x) * (1 - z), x * (1 - z), -(1 - x) * z, -x * z, (1 - x) * z, x * z};
            const double v2[8] = {-(1 - x) * (1 - y), -x * (1 - y), -(1 - x) * y, -x * y, (1 - x) * (1 - y), x * (1 - y), (1 - x) * y, x * y};
            double *B = &Bgp[(size_t)g * nstr * nd];
            for (int q = 0; q < 8; ++q) {
                const double b0 = v0[q] / le[0], b1 = v1[q] / le[1], b2 = v2[q] / le[2];
                if (h == 1) {
                    B[0 * nd + q] = b0, B[1 * nd + q] = b1, B[2 * nd + q] = b2;
                } else {
                    B[0 * nd + 3 * q + 0] = b0, B[1 * nd + 3 * q + 1] = b1, B[2 * nd + 3 * q + 2] = b2;
                    B[3 * nd + 3 * q + 0] = rs * b1, B[3 * nd + 3 * q + 1] = rs * b0;
                    B[4 * nd + 3 * q + 0] = rs * b2, B[4 * nd + 3 * q + 2] = rs * b0;
                    B[5 * nd + 3 * q + 1] = rs * b2, B[5 * nd + 3 * q + 2] = rs * b1;
                }
            }
        }
    }

    // K = sum_gp B^T C B v_e/8 ; F = sum_gp B^T C v_e/8 (nd x nstr: the load vector of a macro gradient g0 is F g0)
    void element_matrices(const double *C, double *K, double *F) const
    {
        std::fill(K, K + nd * nd, 0.0);
        if (F) std::fill(F, F + nd * nstr, 0.0);
        std::vector<double> CB((size_t)nstr * nd);
        for (int g = 0; g < 8; ++g) {
            const double *B = &Bgp[(size_t)g * nstr * nd];
            for (int i = 0; i < nstr; ++i)
                for (int c = 0; c < nd; ++c) {
                    double sum = 0.0;
                    for (int j = 0; j < nstr; ++j) sum += C[i * nstr + j] * B[j * nd + c];
                    CB[i * nd + c] = sum;
                }
            for (int a = 0; a < nd; ++a) {
                for (int c = 0; c < nd; ++c) {
                    double sum = 0.0;
                    for (int i = 0; i < nstr; ++i) sum += B[i * nd + a] * CB[i * nd + c];
                    K[a * nd + c] += sum * ve / 8.0;
                }
                if (F)
                    for (int j = 0; j < nstr; ++j) {
                        double sum = 0.0;
                        for (int i = 0; i < nstr; ++i) sum += B[i * nd + a] * C[i * nstr + j];
                        F[a * nstr + j] += sum * ve / 8.0;
                    }
            }
        }
    }

    void set_phases(int n, const double *C)
    {
        nph = n;
        Cph.assign(C, C + (size_t)n * nstr * nstr);
        Kph.assign((size_t)n * nd * nd, 0.0);
        fgph.assign((size_t)n * nd * nstr, 0.0);
        for (int p = 0; p < n; ++p) element_matrices(C + (size_t)p * nstr * nstr, &Kph[(size_t)p * nd * nd], &fgph[(size_t)p * nd * nstr]);
    }

    // computeFundamentalSolution (solver.h:144-204)
    void build_gamma(const double *Cref)
    {
        std::vector<double> K((size_t)nd * nd), Ker0((size_t)nd * nd);
        element_matrices(Cref, K.data(), nullptr);
        for (int i = 0; i < nd; ++i)      // component-major reordering, matmodel.h:247-251
            for (int j = 0; j < nd; ++j) Ker0[((i % h) * 8 + i / h) * nd + (j % h) * 8 + j / h] = K[i * nd + j];
        const int NG = h * (h + 1) / 2;
        gamma.assign((size_t)nx * ny * kzc * NG, 0.0);
        const double invN = 1.0 / (double)N;
        pool.run(nx, [&](long x0, long x1, int) {
            for (long ix = x0; ix < x1; ++ix)
                for (int iy = 0; iy < ny; ++iy)
                    for (int iz = 0; iz < kzc; ++iz) {
                        double *G = &gamma[(((size_t)ix * ny + iy) * kzc + iz) * NG];
                        if (ix == 0 && iy == 0 && iz == 0) continue;   // Gamma(0) = 0
                        const cplx ex = std::polar(1.0, 2.0 * M_PI * ix / nx), ey = std::polar(1.0, 2.0 * M_PI * iy / ny),
                                   ez = std::polar(1.0, 2.0 * M_PI * iz / nz);
                        const cplx A[8] = {1.0, ex, ey, ex * ey, ez, ex * ez, ez * ey, ex * ey * ez};
                        double AA[8][8];
                        for (int a = 0; a < 8; ++a)
                            for (int b = 0; b < 8; ++b) AA[a][b] = A[a].real() * A[b].real() + A[a].imag() * A[b].imag();
                        double blk[9], inv[9];
                        for (int i = 0; i < h; ++i)
                            for (int j = 0; j < h; ++j) {
                                double sum = 0.0;
                                for (int a = 0; a < 8; ++a)
                                    for (int b = 0; b < 8; ++b) sum += Ker0[(i * 8 + a) * nd + j * 8 + b] * AA[a][b];
                                blk[i * h + j] = sum;
                            }
                        pinv_sym(h, blk, inv);
                        int k = 0;
                        for (int i = 0; i < h; ++i)
                            for (int j = i; j < h; ++j) G[k++] = inv[i * h + j] * invN;
                    }
        });
    }

    // out = sum_e scatter( K_phase (ue - ue[node0]) [+ F_phase g0] )      (solver.h:229-270, one-plane halo by periodic wrap)
    void apply_elements(const double *in, double *out, const double *g0)
    {
        // gather form per node plane is not needed on shared memory: threads own x-slabs of OUTPUT planes and each evaluates the two
        // element planes that touch its planes' nodes only once per plane pair (elements x and x-1), writing to private planes
        std::fill(out, out + N * h, 0.0);
        const int T = pool.nt;
        std::vector<std::vector<double>> carry(T);   // contributions of a thread's last element plane to the next thread's first node plane
        std::vector<long> carry_plane(T, -1);
        pool.run(nx, [&](long x0, long x1, int t) {
            carry[t].assign((size_t)ny * nz * h, 0.0);
            carry_plane[t] = x1 % nx;
            std::vector<double> ue(nd), re(nd);
            for (long ex = x0; ex < x1; ++ex) {
                const long xn = (ex + 1) % nx;
                const bool last = (ex == x1 - 1);
                for (int ey = 0; ey < ny; ++ey) {
                    const int yn = (ey + 1) % ny;
                    for (int ez = 0; ez < nz; ++ez) {
                        const int zn = (ez + 1) % nz;
                        const size_t nodes[8] = {((size_t)ex * ny + ey) * nz + ez, ((size_t)xn * ny + ey) * nz + ez, ((size_t)ex * ny + yn) * nz + ez,
                                                 ((size_t)xn * ny + yn) * nz + ez, ((size_t)ex * ny + ey) * nz + zn, ((size_t)xn * ny + ey) * nz + zn,
                                                 ((size_t)ex * ny + yn) * nz + zn, ((size_t)xn * ny + yn) * nz + zn};
                        for (int a = 0; a < 8; ++a)
                            for (int c = 0; c < h; ++c) ue[h * a + c] = in[nodes[a] * h + c] - in[nodes[0] * h + c];
                        const int ph = ms[nodes[0]];
                        const double *K = &Kph[(size_t)ph * nd * nd];   // symmetric: row j is column j, so the inner loop runs over i
                        for (int i = 0; i < nd; ++i) re[i] = 0.0;
                        for (int j = h; j < nd; ++j) {
                            const double uj = ue[j];
                            const double *Kj = K + (size_t)j * nd;
                            for (int i = 0; i < nd; ++i) re[i] += Kj[i] * uj;
                        }
                        if (g0) {
                            const double *F = &fgph[(size_t)ph * nd * nstr];
                            for (int i = 0; i < nd; ++i)
                                for (int j = 0; j < nstr; ++j) re[i] += F[i * nstr + j] * g0[j];
                        }
                        for (int a = 0; a < 8; ++a) {
                            const bool upper = (a & 1);
                            if (upper && last) {
                                const size_t pn = nodes[a] - (size_t)xn * ny * nz;
                                for (int c = 0; c < h; ++c) carry[t][pn * h + c] += re[h * a + c];
                            } else {
                                for (int c = 0; c < h; ++c) out[nodes[a] * h + c] += re[h * a + c];
                            }
                        }
                    }
                }
            }
        });
        const int used = (int)std::min<long>(T, nx);
        for (int t = 0; t < used; ++t) {   // the r-halo exchange + add of solver.h:262-269
            if (carry_plane[t] < 0) continue;
            double *dst = out + (size_t)carry_plane[t] * ny * nz * h;
            const std::vector<double> &c = carry[t];
            pool.run((long)c.size(), [&](long a, long b, int) {
                for (long i = a; i < b; ++i) dst[i] += c[i];
            });
        }
    }

    // r2c of every z line, then y, then x; Gamma multiply; inverse.  out = Gamma * in (1/N folded into Gamma, solver.h:387-412)
    void convolution(const double *in, double *out)
    {
        const auto t0 = std::chrono::steady_clock::now();
        const int hz = nz / 2;
        // ---- z: real -> half spectrum, per (x, y) line and component
        pool.run((long)nx * ny, [&](long l0, long l1, int) {
            std::vector<cplx> buf(hz);
            for (long l = l0; l < l1; ++l)
                for (int c = 0; c < h; ++c) {
                    const double *src = in + (size_t)l * nz * h + c;
                    for (int k = 0; k < hz; ++k) buf[k] = cplx(src[(size_t)(2 * k) * h], src[(size_t)(2 * k + 1) * h]);
                    fft_lines<1>(buf.data(), hz, pzh.tw.data(), pzh.rev.data(), false);
                    cplx *dst = &spec[(size_t)l * kzc * h + c];
                    for (int k = 0; k <= hz; ++k) {
                        const cplx zk = buf[k % hz], zc = std::conj(buf[(hz - k) % hz]);
                        const cplx e = 0.5 * (zk + zc), o = cplx(0.0, -0.5) * (zk - zc);
                        dst[(size_t)k * h] = e + twz[k] * o;
                    }
                }
        });
        const long W = (long)kzc * h;   // contiguous complex values per (x, y)
        constexpr int VL = 8;
        // ---- y and x: strided lines, VL interleaved columns at a time
        auto strided = [&](bool along_x, bool inverse) {
            const int n = along_x ? nx : ny;
            const long outer = along_x ? ny : nx;
            const Plan &pl = along_x ? px : py;
            const long nblk = (W + VL - 1) / VL;
            pool.run(outer * nblk, [&](long j0, long j1, int) {
                std::vector<cplx> buf((size_t)n * VL);
                for (long j = j0; j < j1; ++j) {
                    const long o = j / nblk, w0 = (j % nblk) * VL;
                    const int wv = (int)std::min<long>(VL, W - w0);
                    for (int i = 0; i < n; ++i) {
                        const size_t base = along_x ? (((size_t)i * ny + o) * W + w0) : (((size_t)o * ny + i) * W + w0);
                        for (int v = 0; v < wv; ++v) buf[(size_t)i * VL + v] = spec[base + v];
                        for (int v = wv; v < VL; ++v) buf[(size_t)i * VL + v] = 0.0;
                    }
                    fft_lines<VL>(buf.data(), n, pl.tw.data(), pl.rev.data(), inverse);
                    for (int i = 0; i < n; ++i) {
                        const size_t base = along_x ? (((size_t)i * ny + o) * W + w0) : (((size_t)o * ny + i) * W + w0);
                        for (int v = 0; v < wv; ++v) spec[base + v] = buf[(size_t)i * VL + v];
                    }
                }
            });
        };
        strided(false, false);
        strided(true, false);
        // ---- Gamma multiply: real symmetric h x h times complex h (solver.h:398-407)
        const int NG = h * (h + 1) / 2;
        pool.run((long)nx * ny * kzc, [&](long f0, long f1, int) {
            for (long f = f0; f < f1; ++f) {
                const double *G = &gamma[(size_t)f * NG];
                cplx *v = &spec[(size_t)f * h];
                if (h == 1) {
                    v[0] *= G[0];
                } else {
                    const cplx a = v[0], b = v[1], c = v[2];
                    v[0] = G[0] * a + G[1] * b + G[2] * c;
                    v[1] = G[1] * a + G[3] * b + G[4] * c;
                    v[2] = G[2] * a + G[4] * b + G[5] * c;
                }
            }
        });
        strided(true, true);
        strided(false, true);
        // ---- z: half spectrum -> real
        pool.run((long)nx * ny, [&](long l0, long l1, int) {
            std::vector<cplx> buf(hz);
            for (long l = l0; l < l1; ++l)
                for (int c = 0; c < h; ++c) {
                    const cplx *src = &spec[(size_t)l * kzc * h + c];
                    for (int k = 0; k < hz; ++k) {
                        const cplx xk = src[(size_t)k * h], xc = std::conj(src[(size_t)(hz - k) * h]);
                        const cplx e = xk + xc, o = (xk - xc) * std::conj(twz[k]);
                        buf[k] = e + cplx(0.0, 1.0) * o;    // = 2 z_k of the packed half-length sequence
                    }
                    fft_lines<1>(buf.data(), hz, pzh.tw.data(), pzh.rev.data(), true);
                    double *dst = out + (size_t)l * nz * h + c;
                    for (int k = 0; k < hz; ++k) {
                        dst[(size_t)(2 * k) * h] = buf[k].real();   // FFTW c2r is unscaled: sum over all nz frequencies
                        dst[(size_t)(2 * k + 1) * h] = buf[k].imag();
                    }
                }
        });
        fft_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }

    double dot(const double *a, const double *b) const
    {
        std::vector<double> part(pool.nt, 0.0);
        pool.run((long)(N * h), [&](long i0, long i1, int t) {
            double sum = 0.0;
            for (long i = i0; i < i1; ++i) sum += a[i] * b[i];
            part[t] = sum;
        });
        double sum = 0.0;
        for (double p : part) sum += p;
        return sum;
    }

    double norm(const double *a, int measure) const   // 0 L1, 1 L2, 2 Linf
    {
        std::vector<double> part(pool.nt, 0.0);
        pool.run((long)(N * h), [&](long i0, long i1, int t) {
            double sum = 0.0;
            for (long i = i0; i < i1; ++i) {
                const double v = std::fabs(a[i]);
                if (measure == 0) sum += v;
                else if (measure == 1) sum += v * v;
                else sum = std::max(sum, v);
            }
            part[t] = sum;
        });
        double sum = 0.0;
        for (double p : part) sum = (measure == 2) ? std::max(sum, p) : sum + p;
        return measure == 1 ? std::sqrt(sum) : sum;
    }

    // get_homogenized_stress for linear phases: mean over elements and Gauss points of C (B ue + g0)   (solver.h:707-737)
    void homogenized_stress(const double *g0, double *out) const
    {
        std::vector<std::vector<double>> part(pool.nt, std::vector<double>(nstr, 0.0));
        std::vector<double> Bavg((size_t)nstr * nd, 0.0);
        for (int g = 0; g < 8; ++g)
            for (int i = 0; i < nstr * nd; ++i) Bavg[i] += Bgp[(size_t)g * nstr * nd + i] / 8.0;
        pool.run(nx, [&](long x0, long x1, int t) {
            std::vector<double> ue(nd), eps(nstr);
            for (long ex = x0; ex < x1; ++ex) {
                const long xn = (ex + 1) % nx;
                for (int ey = 0; ey < ny; ++ey) {
                    const int yn = (ey + 1) % ny;
                    for (int ez = 0; ez < nz; ++ez) {
                        const int zn = (ez + 1) % nz;
                        const size_t nodes[8] = {((size_t)ex * ny + ey) * nz + ez, ((size_t)xn * ny + ey) * nz + ez, ((size_t)ex * ny + yn) * nz + ez,
                                                 ((size_t)xn * ny + yn) * nz + ez, ((size_t)ex * ny + ey) * nz + zn, ((size_t)xn * ny + ey) * nz + zn,
                                                 ((size_t)ex * ny + yn) * nz + zn, ((size_t)xn * ny + yn) * nz + zn};
                        for (int a = 0; a < 8; ++a)
                            for (int c = 0; c < h; ++c) ue[h * a + c] = u[nodes[a] * h + c];
                        for (int i = 0; i < nstr; ++i) {
                            double sum = g0[i];
                            for (int j = 0; j < nd; ++j) sum += Bavg[i * nd + j] * ue[j];
                            eps[i] = sum;
                        }
                        const double *C = &Cph[(size_t)ms[nodes[0]] * nstr * nstr];
                        for (int i = 0; i < nstr; ++i) {
                            double sum = 0.0;
                            for (int j = 0; j < nstr; ++j) sum += C[i * nstr + j] * eps[j];
                            part[t][i] += sum;
                        }
                    }
                }
            }
        });
        for (int i = 0; i < nstr; ++i) {
            double sum = 0.0;
            for (auto &p : part) sum += p[i];
            out[i] = sum / (double)N;
        }
    }
};

}  // namespace

extern "C" {

void *fcpu_create(int nx, int ny, int nz, const double *L, int howmany, int nthreads)
{
    for (int n : {nx, ny, nz})
        if (n < 4 || (n & (n - 1))) return nullptr;   // radix-2 FFT
    if (howmany != 1 && howmany != 3) return nullptr;
    return new Cpu(nx, ny, nz, L, howmany, nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency());
}
void fcpu_destroy(void *p) { delete (Cpu *)p; }
int fcpu_threads(void *p) { return ((Cpu *)p)->pool.nt; }
void fcpu_set_microstructure(void *p, const uint16_t *ms)
{
    Cpu *c = (Cpu *)p;
    c->ms.assign(ms, ms + c->N);
}
void fcpu_set_phases(void *p, int n, const double *C) { ((Cpu *)p)->set_phases(n, C); }
void fcpu_set_reference(void *p, const double *Cref) { ((Cpu *)p)->build_gamma(Cref); }
void fcpu_get_u(void *p, double *out)
{
    Cpu *c = (Cpu *)p;
    std::memcpy(out, c->u.data(), sizeof(double) * c->N * c->h);
}
void fcpu_zero_u(void *p)
{
    Cpu *c = (Cpu *)p;
    std::fill(c->u.begin(), c->u.end(), 0.0);
}
void fcpu_convolution(void *p, const double *in, double *out) { ((Cpu *)p)->convolution(in, out); }
void fcpu_apply_linear(void *p, const double *in, double *out) { ((Cpu *)p)->apply_elements(in, out, nullptr); }

// SolverCG::internalSolve, linear branch.  times[0] = seconds of the iteration loop (after the initial residual), times[1] = seconds
// inside convolution() (solver.h:293 "FFT Time").  Returns the iteration count.
int fcpu_solve_cg(void *p, const double *g0, int n_it, double tol, int measure, double *err_hist, double *sigma_out, double *times)
{
    Cpu &c = *(Cpu *)p;
    const size_t n = c.N * c.h;
    std::fill(c.s.begin(), c.s.end(), 0.0);
    std::fill(c.d.begin(), c.d.end(), 0.0);
    c.apply_elements(c.u.data(), c.r.data(), g0);
    int iter = 0;
    double err = c.norm(c.r.data(), measure);
    if (err_hist) err_hist[0] = err;
    double delta = 1.0;
    c.fft_seconds = 0.0;
    const auto t0 = std::chrono::steady_clock::now();
    while (iter < n_it && err > tol) {
        const double deltamid = c.dot(c.r.data(), c.s.data());
        c.convolution(c.r.data(), c.s.data());
        c.pool.run((long)n, [&](long a, long b, int) {
            for (long i = a; i < b; ++i) c.s[i] = -c.s[i];
        });
        const double delta0 = delta;
        delta = c.dot(c.r.data(), c.s.data());
        const double beta = std::fmax(0.0, (delta - deltamid) / delta0);
        c.pool.run((long)n, [&](long a, long b, int) {
            for (long i = a; i < b; ++i) c.d[i] = c.s[i] + beta * c.d[i];
        });
        c.apply_elements(c.d.data(), c.rnew.data(), nullptr);
        const double alpha = delta / c.dot(c.d.data(), c.rnew.data());
        c.pool.run((long)n, [&](long a, long b, int) {
            for (long i = a; i < b; ++i) {
                c.r[i] -= alpha * c.rnew[i];
                c.u[i] -= alpha * c.d[i];
            }
        });
        ++iter;
        err = c.norm(c.r.data(), measure);
        if (err_hist) err_hist[iter] = err;
    }
    if (times) {
        times[0] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        times[1] = c.fft_seconds;
    }
    if (sigma_out) c.homogenized_stress(g0, sigma_out);
    return iter;
}

}  // extern "C"
