/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the
 * product path (maicos_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * Plain-C restatement of the reference's reciprocal-space structure factor loop,
 *   /root/reference/src/maicos/lib/_cmath.pyx:21-126  (cpdef structure_factor).
 * Parity status: PINNED -- checked against (i) the reference kernel itself compiled
 * from /root/reference into oracle/_ref (tests/test_oracle.py, bit-level mask equality
 * and <=1e-12 relative on S) and (ii) the reference's golden vectors
 * tests/lib/test_math.py:183-209 (six-vector theta table) via tests/golden/.
 *
 * Build: see oracle/Makefile (gcc -O3 -ffast-math -fopenmp, the reference's own flags,
 * setup.py:100-110).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* _cmath.pyx:93-96 -- q_factor[i] = 2*pi/L[i]; maxn[i] = (int)ceil(qmax / (float)q_factor[i])
 * NB: the divisor is cast to float32 first; qmax/(float) is then evaluated in double. */
void oracle_sfactor_maxn(const double *dimensions, double qmax, int32_t *maxn) {
    for (int i = 0; i < 3; ++i) {
        double q_factor = 2 * M_PI / dimensions[i];
        maxn[i] = (int32_t)ceil(qmax / (double)(float)q_factor);
    }
}

/* _cmath.pyx:98-124, restricted to the h-slab h0 <= h < h1 of the reciprocal cube (the full call is the slab
 * [0, maxn0)); cells outside the slab stay zero.  positions is [n_atoms][3] with arbitrary element strides (in
 * doubles), outputs are dense C-contiguous [maxn0][maxn1][maxn2], zero where rejected.  Cells are independent, so the
 * (h, k) loops are shared among the threads (the reference's prange runs over h alone, :101). */
void oracle_structure_factor_slab(const double *positions, ptrdiff_t stride_atom, ptrdiff_t stride_xyz,
                                  ptrdiff_t n_atoms, const double *dimensions, double qmin,
                                  double qmax, double thetamin, double thetamax,
                                  const double *weights, int32_t h0, int32_t h1,
                                  double *scattering_vectors, double *structure_factors) {
    int32_t maxn[3];
    double q_factor[3];
    oracle_sfactor_maxn(dimensions, qmax, maxn);
    for (int i = 0; i < 3; ++i) q_factor[i] = 2 * M_PI / dimensions[i];
    size_t cells = (size_t)maxn[0] * maxn[1] * maxn[2];
    memset(scattering_vectors, 0, cells * sizeof(double));
    memset(structure_factors, 0, cells * sizeof(double));
    if (h0 < 0) h0 = 0;
    if (h1 > maxn[0]) h1 = maxn[0];

#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (int h = h0; h < h1; ++h) {
        for (int k = 0; k < maxn[1]; ++k) {
            double qx = h * q_factor[0];
            double qy = k * q_factor[1];
            for (int l = 0; l < maxn[2]; ++l) {
                if (h + k + l == 0) continue;                      /* :108 */
                double qz = l * q_factor[2];
                double qrr = sqrt(qx * qx + qy * qy + qz * qz);   /* :110 */
                double theta = acos(qz / qrr);                    /* :111 */
                if (!(qrr >= qmin && qrr <= qmax && theta >= thetamin && theta <= thetamax))
                    continue;                                      /* :113-114 */
                size_t idx = ((size_t)h * maxn[1] + k) * maxn[2] + l;
                scattering_vectors[idx] = qrr;                     /* :115 */
                double s = 0.0, c = 0.0;
                for (ptrdiff_t j = 0; j < n_atoms; ++j) {          /* :119-122 */
                    const double *p = positions + j * stride_atom;
                    double qdotr = p[0] * qx + p[stride_xyz] * qy + p[2 * stride_xyz] * qz;
                    s += weights[j] * sin(qdotr);
                    c += weights[j] * cos(qdotr);
                }
                structure_factors[idx] += s * s + c * c;           /* :124 */
            }
        }
    }
}

void oracle_structure_factor(const double *positions, ptrdiff_t stride_atom, ptrdiff_t stride_xyz,
                             ptrdiff_t n_atoms, const double *dimensions, double qmin,
                             double qmax, double thetamin, double thetamax,
                             const double *weights, double *scattering_vectors,
                             double *structure_factors) {
    oracle_structure_factor_slab(positions, stride_atom, stride_xyz, n_atoms, dimensions, qmin, qmax, thetamin,
                                 thetamax, weights, 0, INT32_MAX, scattering_vectors, structure_factors);
}

/* OpenMP pool size of this process (bench.py: torchrun exports OMP_NUM_THREADS=1 to its workers). */
#ifdef _OPENMP
#include <omp.h>
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_get_threads(void) { return omp_get_max_threads(); }
#else
void oracle_set_threads(int n) { (void)n; }
int oracle_get_threads(void) { return 1; }
#endif
