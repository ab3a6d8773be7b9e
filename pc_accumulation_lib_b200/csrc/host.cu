// host.cu — host-side helpers of libpcacc that sit on the end-to-end path of the
// reference-facing API: cropping / pre-processing of the trajectory polylines that
// accompany every BEV.  No device code.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

// ---------------------------------------------------------------------------
// host-side trajectory pre-processing (same double arithmetic as the reference's Python)
// ---------------------------------------------------------------------------
static inline bool in_box(double x, double y, double h) { return (-h < x && x < h) && (-h < y && y < h); }

// crop one polyline (n rows of 3 doubles) against the open box (-view/2, view/2)^2
static int crop_polyline(const double *traj, int n, double view, double thresh, double *out) {
    const double h = 0.5 * view;
    int m = 0;
    for (int k = 0; k + 1 < n; k++) {
        double x0 = traj[3 * k], y0 = traj[3 * k + 1], x1 = traj[3 * k + 3], y1 = traj[3 * k + 4];
        const double z0 = traj[3 * k + 2];
        const bool a_in = in_box(x0, y0, h), b_in = in_box(x1, y1, h);
        if (a_in) {
            out[3 * m] = x0;
            out[3 * m + 1] = y0;
            out[3 * m + 2] = z0;
            m++;
        }
        if (a_in != b_in) {
            double xm = 0.0, ym = 0.0, moved = INFINITY;
            while (moved > thresh) {
                xm = 0.5 * (x0 + x1);
                ym = 0.5 * (y0 + y1);
                const bool first_in = in_box(x0, y0, h), mid_in = in_box(xm, ym, h);
                if (mid_in == first_in) {   // midpoint on the side of end 0: replace end 0
                    const double dx = xm - x0, dy = ym - y0;
                    moved = sqrt(dx * dx + dy * dy);
                    x0 = xm;
                    y0 = ym;
                } else {
                    const double dx = xm - x1, dy = ym - y1;
                    moved = sqrt(dx * dx + dy * dy);
                    x1 = xm;
                    y1 = ym;
                }
            }
            out[3 * m] = xm;
            out[3 * m + 1] = ym;
            out[3 * m + 2] = z0;
            m++;
        }
    }
    return m;
}

extern "C" int pcacc_crop_trajectory(const double *traj, int n, double view, double thresh, double *out,
                                     int *n_out) {
    if (!traj || !out || !n_out || n < 0) return PCACC_ERR_ARG;
    *n_out = crop_polyline(traj, n, view, thresh, out);
    return PCACC_OK;
}

extern "C" int pcacc_preprocess_trajectories(const double *pts, const int32_t *traj_off, int n_traj,
                                             const double *variants, int n_var, int P, double thresh,
                                             double *out, int32_t *out_cnt) {
    if (n_traj < 0 || n_var < 0 || P <= 0 || (n_traj && (!traj_off || !out_cnt)) ||
        (n_var && !variants))
        return PCACC_ERR_ARG;
    if (n_traj == 0 || n_var == 0) return PCACC_OK;
    const int64_t N = traj_off[n_traj];
    if (N && (!pts || !out)) return PCACC_ERR_ARG;
    std::vector<double> tmp;
    for (int v = 0; v < n_var; v++) {
        const double *R = variants + 12 * (size_t)v;
        const double dx = R[9], dy = R[10], view = R[11];
        if (!(view > 0.0)) return PCACC_ERR_ARG;
        const double dP = (double)P, hP = 0.5 * dP;
        for (int t = 0; t < n_traj; t++) {
            const int b = traj_off[t], n = traj_off[t + 1] - b;
            if (n < 0) return PCACC_ERR_ARG;
            int m = 0;
            double *o = out + 3 * ((size_t)v * 2 * (size_t)N + 2 * (size_t)b);
            if (n >= 2) {
                tmp.resize(3 * (size_t)n);
                for (int i = 0; i < n; i++) {
                    const double *p = pts + 3 * (size_t)(b + i);
                    // np.matmul(R, p.T).T: one fused multiply-add chain over k per element
                    for (int r = 0; r < 3; r++)
                        tmp[3 * i + r] = fma(R[3 * r + 2], p[2], fma(R[3 * r + 1], p[1], R[3 * r] * p[0]));
                    tmp[3 * i] += dx;
                    tmp[3 * i + 1] += dy;
                }
                m = crop_polyline(tmp.data(), n, view, thresh, o);
                for (int i = 0; i < m; i++) {   // pos2grid: div, mul, add rounded separately
                    o[3 * i] = floor(o[3 * i] / view * dP + hP);
                    o[3 * i + 1] = floor(o[3 * i + 1] / view * dP + hP);
                }
            }
            out_cnt[(size_t)v * n_traj + t] = m;
        }
    }
    return PCACC_OK;
}
