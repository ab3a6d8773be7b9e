/* TEST INFRASTRUCTURE — CPU oracle helper (not product code).
 *
 * numpy's float64 matmul for the (3,4)@(4,N), (4,4)@(4,N) and (3,3)@(3,N)
 * products on the reference's hot path (sem_pc_accum.py:362,
 * sem_pc_accum.py:179, datasets/nuscenes_utils.py:58,
 * bev_generator/bev_generator.py:227) is, for N >= 2, bit-identical to a
 * sequential fused-multiply-add chain over k (re-checked by
 * oracle/validate_against_reference.py and tests/test_oracle_cpu.py).  numpy
 * has no fma, and which OpenBLAS kernel runs depends on the host CPU, so the
 * oracle states that arithmetic once, in C, with libm's correctly rounded
 * fma(); the CUDA kernels use __fma_rn in the same order.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>

/* out[i,r] = (((M[r,0]*x) + M[r,1]*y) + M[r,2]*z) + M[r,3]*1  with every
 * "+ a*b" fused; M is row-major with leading dimension ld (3x4 or 4x4). */
void orc_affine_f64(const double *M, int ld, const double *pts, int64_t n,
                    int64_t pts_stride, double *out)
{
    for (int64_t i = 0; i < n; i++) {
        const double *p = pts + i * pts_stride;
        for (int r = 0; r < 3; r++) {
            const double *m = M + r * ld;
            double a = m[0] * p[0];
            a = fma(m[1], p[1], a);
            a = fma(m[2], p[2], a);
            a = fma(m[3], 1.0, a);
            out[3 * i + r] = a;
        }
    }
}

/* same with float32 points promoted to float64 first (np.concatenate of an
 * f32 cloud with an f64 ones column, sem_pc_accum.py:358) */
void orc_affine_f32(const double *M, int ld, const float *pts, int64_t n,
                    int64_t pts_stride, double *out)
{
    for (int64_t i = 0; i < n; i++) {
        const float *p = pts + i * pts_stride;
        for (int r = 0; r < 3; r++) {
            const double *m = M + r * ld;
            double a = m[0] * (double)p[0];
            a = fma(m[1], (double)p[1], a);
            a = fma(m[2], (double)p[2], a);
            a = fma(m[3], 1.0, a);
            out[3 * i + r] = a;
        }
    }
}

/* out[i,r] = fma chain over k = 0..2 of R[r,k]*p[k]  (3x3 rotation) */
void orc_rot33_f64(const double *R, const double *pts, int64_t n,
                   int64_t pts_stride, double *out)
{
    for (int64_t i = 0; i < n; i++) {
        const double *p = pts + i * pts_stride;
        for (int r = 0; r < 3; r++) {
            const double *m = R + r * 3;
            double a = m[0] * p[0];
            a = fma(m[1], p[1], a);
            a = fma(m[2], p[2], a);
            out[3 * i + r] = a;
        }
    }
}
