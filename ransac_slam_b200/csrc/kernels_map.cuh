// kernels_map.cuh -- Map state surgery on the device: the covariance never leaves HBM while the map changes.
//
//   Map::delete_a_feature                         (src/Map.cpp:69-104)    remove a 3/6-wide block of rows and columns
//   Map::inversedepth_2_cartesian                 (src/Map.cpp:105-196)   linearity test, then P <- J P J^T with J = blockdiag(I, J_3x6, I)
//   Map::add_a_feature_covariance_inverse_depth   (src/Map.cpp:339-400) + ExtendKF::hinv (src/ExtendKF.cpp:236-265)
//
// All three are the same shape of work:  P' = T P T^T (+ E),  x' = T x,  where T copies most rows and forms a few "special" rows as
// small linear combinations of a contiguous range of old rows (none for delete, 3 rows from 6 for the conversion, 6 rows from the 13
// camera rows for a new feature).  The reference multiplies dense n x n matrices for this (J_all * P * J_all^T, src/Map.cpp:171-188);
// here one out-of-place pass writes P' into a scratch covariance (thread per output entry, coalesced down the columns) and the
// handle swaps the two buffers: 16 n^2 bytes of HBM traffic, no flops to speak of.
#pragma once
#include "common.cuh"

namespace rslam {

struct MapXform {
    int mode;      // 0 delete, 1 inverse depth -> cartesian, 2 add inverse-depth feature
    int n_old, n_new;
    int cut_at;    // first old row that is dropped (delete / convert); n_old for add
    int cut_cnt;   // number of old rows dropped (3 or 6 / 6 / 0)
    int sp_at;     // first special row of the new state (n_new if none)
    int sp_cnt;    // special rows: 0 / 3 / 6
    int src_at;    // first old row the special rows combine
    int src_cnt;   // 6 / 13
};

// scratch layout (doubles) written by the prep kernels: coefficient matrix C[sp_cnt][src_cnt] row-major at 0, additive block
// E[6][6] at 96, special entries of x' at 132
constexpr int kMapC = 0, kMapE = 96, kMapX = 132, kMapScratch = 160;

__device__ __forceinline__ int map_old_row(const MapXform& xf, int r) {  // new copy-row -> old row
    if (r < xf.cut_at) return r;
    return r + xf.cut_cnt - xf.sp_cnt;  // delete: r + cut_cnt ; convert: r + 3 ; add: never reached for r >= n_old (special)
}

// linearity index of every inverse-depth feature (src/Map.cpp:113-148); the first feature (feature order) below the threshold is
// the one the reference converts (it returns after one conversion).  result[0] = that feature index or INT_MAX.
__global__ void k_map_linearity(DevFilter* Fs, int b, double threshold, int* result) {
    const DevFilter& F = Fs[b];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F.N || F.ftype[i] != 0) return;
    const int ip = F.foff[i];
    const double* X = F.x_kk;
    const double std_rho = sqrt(F.P[(ip + 5) + (size_t)(ip + 5) * F.ldp]);
    const double rho = X[ip + 5];
    const double std_d = std_rho / (rho * rho);
    const double theta = X[ip + 3], phi = X[ip + 4];
    const double mi[3] = {cos(phi) * sin(theta), -sin(phi), cos(phi) * cos(theta)};
    double a = 0, n1 = 0, n2 = 0;
    for (int k = 0; k < 3; k++) {
        const double xo = X[ip + k] + mi[k] / rho;  // inversedepth2cartesian (src/ExtendKF.cpp:137-152)
        const double d1 = xo - X[ip + k], d2 = xo - X[k];
        a += d1 * d2;
        n1 += d1 * d1;
        n2 += d2 * d2;
    }
    const double d_c2p = sqrt(n2);
    const double cos_alpha = a / (sqrt(n1) * sqrt(n2));
    const double linearity_index = 4 * std_d * cos_alpha / d_c2p;
    if (linearity_index < threshold) atomicMin(result, i);
}

// coefficients of the conversion of the feature at state offset ip (src/Map.cpp:150-169): J = [I3, dm/dtheta / rho, dm/dphi / rho, -m / rho^2]
__global__ void k_map_prep_convert(DevFilter* Fs, int b, int ip, double* scratch) {
    const DevFilter& F = Fs[b];
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double* X = F.x_kk;
    const double theta = X[ip + 3], phi = X[ip + 4], rho = X[ip + 5];
    const double mi[3] = {cos(phi) * sin(theta), -sin(phi), cos(phi) * cos(theta)};
    const double dmt[3] = {cos(phi) * cos(theta), 0.0, -cos(phi) * sin(theta)};
    const double dmp[3] = {-sin(phi) * sin(theta), -cos(phi), -sin(phi) * cos(theta)};
    double* C = scratch + kMapC;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) C[r * 6 + c] = (r == c) ? 1.0 : 0.0;
        C[r * 6 + 3] = (1 / rho) * dmt[r];
        C[r * 6 + 4] = (1 / rho) * dmp[r];
        C[r * 6 + 5] = -mi[r] / (rho * rho);
        scratch[kMapX + r] = X[ip + r] + mi[r] / rho;
    }
    for (int e = 0; e < 36; e++) scratch[kMapE + e] = 0.0;
}

// new inverse-depth feature from the distorted pixel uvd at the current camera state: hinv (src/ExtendKF.cpp:236-265), the Jacobians
// dy/dxv (6 x 13) and dy/dhd (6 x 3) and E = dy/dhd * diag(std_z^2, std_z^2, std_rho^2) * dy/dhd^T (src/Map.cpp:339-386); also fills
// the per-feature record of Map::initialize_a_features (src/Map.cpp:286-311): 41 x 41 patch cut from the current image, pose at
// initialisation, pixel.
__global__ void k_map_prep_add(DevFilter* Fs, int b, CamDev cam, double ud, double vd, double initial_rho, double std_pxl, double std_rho,
                               double* scratch) {
    DevFilter& F = Fs[b];
    const int i_new = F.N;  // the descriptor still holds the old N
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double* Xv = F.x_kk;
        double R[9];
        q2r_dev(Xv + 3, R);
        // undistort_fm (src/ExtendKF.cpp:266-285)
        const double xd = (ud - cam.Cx) * cam.dx, yd = (vd - cam.Cy) * cam.dy;
        const double rd = sqrt(xd * xd + yd * yd);
        const double D = 1 + cam.k1 * rd * rd + cam.k2 * rd * rd * rd * rd;
        const double uu = xd * D / cam.dx + cam.Cx, vu = yd * D / cam.dy + cam.Cy;
        const double gc[3] = {-(cam.Cx - uu) / cam.fku, -(cam.Cy - vu) / cam.fkv, 1.0};
        double gw[3];
        for (int k = 0; k < 3; k++) gw[k] = R[3 * k] * gc[0] + R[3 * k + 1] * gc[1] + R[3 * k + 2] * gc[2];
        const double Xw = gw[0], Yw = gw[1], Zw = gw[2];
        double* xs = scratch + kMapX;
        xs[0] = Xv[0];
        xs[1] = Xv[1];
        xs[2] = Xv[2];
        xs[3] = atan2(Xw, Zw);
        xs[4] = atan2(-Yw, sqrt(Xw * Xw + Zw * Zw));
        xs[5] = initial_rho;
        double dgw_dq[12];
        dRq_times_a_by_dq_dev(Xv + 3, gc, dgw_dq);
        const double xz2 = Xw * Xw + Zw * Zw, n2 = xz2 + Yw * Yw, sxz = sqrt(xz2);
        const double dth[3] = {Zw / xz2, 0.0, -Xw / xz2};
        const double dph[3] = {(Xw * Yw) / (n2 * sxz), -sxz / n2, (Zw * Yw) / (n2 * sxz)};
        double* C = scratch + kMapC;  // dy_dxv, 6 x 13 row-major
        for (int e = 0; e < 78; e++) C[e] = 0.0;
        C[0 * 13 + 0] = C[1 * 13 + 1] = C[2 * 13 + 2] = 1.0;
        for (int c = 0; c < 4; c++) {
            C[3 * 13 + 3 + c] = dth[0] * dgw_dq[c] + dth[1] * dgw_dq[4 + c] + dth[2] * dgw_dq[8 + c];
            C[4 * 13 + 3 + c] = dph[0] * dgw_dq[c] + dph[1] * dgw_dq[4 + c] + dph[2] * dgw_dq[8 + c];
        }
        // dyprima_dhd = dyprima_dgw * R_wc * dgc_dhu * dhu_dhd (left to right, src/Map.cpp:371)
        double J[4];
        jacob_undistort_dev(cam, ud, vd, J);
        double tR[2][3];  // rows theta, phi of dyprima_dgw * R_wc
        for (int c = 0; c < 3; c++) {
            tR[0][c] = dth[0] * R[c] + dth[1] * R[3 + c] + dth[2] * R[6 + c];
            tR[1][c] = dph[0] * R[c] + dph[1] * R[3 + c] + dph[2] * R[6 + c];
        }
        // * dgc_dhu.  The reference comma-initialises the 3 x 2 matrix with "1/fku, 0, 0, 0, 1/fkv, 0" (src/Map.cpp:375-379); Eigen
        // fills row by row, so what it multiplies by is [1/fku 0; 0 0; 1/fkv 0], not the [1/fku 0; 0 1/fkv; 0 0] of the MATLAB
        // original (quirk Q16, found by compiling and running the reference's own Map.cpp).  Reproduced: the new feature's
        // (theta, phi) covariance block is rank one.
        double tg[2][2];
        for (int r = 0; r < 2; r++) {
            tg[r][0] = tR[r][0] * (1 / cam.fku) + tR[r][2] * (1 / cam.fkv);
            tg[r][1] = 0.0;
        }
        double dy_dhd[6][3];
        for (int r = 0; r < 6; r++)
            for (int c = 0; c < 3; c++) dy_dhd[r][c] = 0.0;
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 2; c++) dy_dhd[3 + r][c] = tg[r][0] * J[c] + tg[r][1] * J[2 + c];
        dy_dhd[5][2] = 1.0;
        const double pad[3] = {std_pxl * std_pxl, std_pxl * std_pxl, std_rho * std_rho};
        for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++) {
                double s = 0;
                for (int k = 0; k < 3; k++) s += (dy_dhd[r][k] * pad[k]) * dy_dhd[c][k];
                scratch[kMapE + r * 6 + c] = s;
            }
        // per-feature record
        double* ip = F.init_pose + (size_t)i_new * 14;
        ip[0] = Xv[0];
        ip[1] = Xv[1];
        ip[2] = Xv[2];
        for (int e = 0; e < 9; e++) ip[3 + e] = R[e];
        ip[12] = ud;
        ip[13] = vd;
        F.ftype[i_new] = 0;
        F.foff[i_new] = F.n;
        F.times_predicted[i_new] = 0;
        F.times_measured[i_new] = 0;
        F.has_h[i_new] = F.ic[i_new] = F.li[i_new] = F.hi[i_new] = 0;
        F.last_id[i_new] = i_new;
    }
    // 41 x 41 patch around the pixel (cv::Range truncates the double bounds toward zero, src/Map.cpp:286)
    const int u0 = (int)(ud - 20), v0 = (int)(vd - 20);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 1681; e += gridDim.x * blockDim.x) {
        const int r = e / 41, c = e % 41;
        const int gy = v0 + r, gx = u0 + c;
        unsigned char v = 0;
        if (F.image && gx >= 0 && gx < F.img_cols && gy >= 0 && gy < F.img_rows) v = F.image[(size_t)gy * F.img_stride + gx];
        F.patch_init[(size_t)i_new * 1681 + e] = v;
    }
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < kPatchPix; e += gridDim.x * blockDim.x) F.patch[(size_t)i_new * kPatchPix + e] = 0.f;
}

// P' = T P T^T + E and x' = T x, out of place.  grid (ceil(n_new / 256), n_new): thread per entry, blockIdx.y = column.
__global__ void __launch_bounds__(256) k_map_xform(DevFilter* Fs, int b, MapXform xf, const double* scratch, double* Pdst, double* xdst) {
    const DevFilter& F = Fs[b];
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r >= xf.n_new) return;
    const int ld = F.ldp;
    const double* P = F.P;
    const double* C = scratch + kMapC;
    const bool rs = r >= xf.sp_at && r < xf.sp_at + xf.sp_cnt;
    const bool cs = c >= xf.sp_at && c < xf.sp_at + xf.sp_cnt;
    double v;
    if (!rs && !cs) {
        v = P[map_old_row(xf, r) + (size_t)map_old_row(xf, c) * ld];
    } else if (rs && !cs) {
        const double* col = P + (size_t)map_old_row(xf, c) * ld + xf.src_at;
        const double* cr = C + (r - xf.sp_at) * xf.src_cnt;
        v = 0;
        for (int k = 0; k < xf.src_cnt; k++) v += cr[k] * col[k];
    } else if (!rs && cs) {
        const double* row = P + map_old_row(xf, r) + (size_t)xf.src_at * ld;
        const double* cc = C + (c - xf.sp_at) * xf.src_cnt;
        v = 0;
        for (int k = 0; k < xf.src_cnt; k++) v += row[(size_t)k * ld] * cc[k];
    } else {
        // evaluated for the lower-triangle position and mirrored, so that P stays exactly symmetric (the library's invariant)
        const int rr = r > c ? r : c, c2 = r > c ? c : r;
        const double* cr = C + (rr - xf.sp_at) * xf.src_cnt;
        const double* cc = C + (c2 - xf.sp_at) * xf.src_cnt;
        v = 0;
        for (int k = 0; k < xf.src_cnt; k++) {  // (C P) C^T, left to right like the reference's products
            double t = 0;
            for (int l = 0; l < xf.src_cnt; l++) t += cr[l] * P[(xf.src_at + l) + (size_t)(xf.src_at + k) * ld];
            v += t * cc[k];
        }
        v += scratch[kMapE + (rr - xf.sp_at) * 6 + (c2 - xf.sp_at)];
    }
    Pdst[r + (size_t)c * ld] = v;
    if (c == 0) xdst[r] = rs ? scratch[kMapX + (r - xf.sp_at)] : F.x_kk[map_old_row(xf, r)];
}

// ---- feature initialisation: cv::FAST (TYPE_9_16, non-maximum suppression) + the "features in the box" test ----
// cv::FAST as the reference calls it (src/Map.cpp:324-338: cv::FAST(im, keypoints, 100, true)), from the published algorithm
// (OpenCV 3.2 modules/features2d/src/fast.cpp, fast_score.cpp): corner iff 9 contiguous pixels of the radius-3 Bresenham circle are
// all darker than v - t or all brighter than v + t; score = the largest t for which that holds; a corner survives iff its score is
// strictly greater than its 8 neighbours' (non-corners score 0); the 3-pixel border is never tested; output row by row, left to right.
// pass 1: score of every pixel of the window (x0, y0, w, h) of the filter's image; thread per pixel
__global__ void __launch_bounds__(256) k_fast9_score(DevFilter* Fs, int b, int x0, int y0, int w, int h, int threshold, int* score) {
    const DevFilter& F = Fs[b];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= w * h) return;
    const int x = e % w, y = e / w;
    int out = 0;
    if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {
        const unsigned char* img = F.image + (size_t)(y0 + y) * F.img_stride + (x0 + x);
        const int v = img[0];
        const int RX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
        const int RY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
        int d[16];
#pragma unroll 1
        for (int k = 0; k < 16; k++) d[k] = v - (int)img[RY[k] * F.img_stride + RX[k]];
        // NOTE: written with plain compare-and-assign loops that are NOT unrolled on purpose.  The obvious fully unrolled
        // min()/max() form of this arc search is miscompiled by nvcc 12.9 for sm_100a (the fused three-input VIMNMX3 chain returns
        // wrong strengths, e.g. 36 instead of 117; reproduced in isolation by tools/ubench/fast_dbg.cu), so keep it this way.
        int best = -256;
#pragma unroll 1
        for (int s0 = 0; s0 < 16; s0++) {
            int a = 256, bb = 256;
#pragma unroll 1
            for (int j = 0; j < 9; j++) {
                const int t = d[(s0 + j) & 15];
                if (t < a) a = t;
                if (-t < bb) bb = -t;
            }
            const int m = a > bb ? a : bb;
            if (m > best) best = m;
        }
        if (best > threshold) out = best - 1;
    }
    score[e] = out;
}

// pass 2: non-maximum suppression and ORDERED compaction (row-major) by one CTA: res[0] = number of keypoints, xy[2 i] = (x, y)
__global__ void __launch_bounds__(1024) k_fast9_nms(const int* score, int w, int h, int max_kp, int* res, int* xy) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int e0 = 0; e0 < w * h; e0 += 1024) {
        const int e = e0 + threadIdx.x;
        bool keep = false;
        int x = 0, y = 0;
        if (e < w * h) {
            x = e % w;
            y = e / w;
            const int sc = score[e];
            if (sc > 0) {  // inside the tested area by construction, so all 8 neighbours exist
                keep = sc > score[e - 1] && sc > score[e + 1] && sc > score[e - w - 1] && sc > score[e - w] && sc > score[e - w + 1] &&
                       sc > score[e + w - 1] && sc > score[e + w] && sc > score[e + w + 1];
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int k = 0; k < 32; k++) {
            const int c = s_warp[k];
            if (k < wid) before += c;
            total += c;
        }
        if (keep) {
            const int pos = s_base + before + __popc(bal & ((1u << lane) - 1));
            if (pos < max_kp) {
                xy[2 * pos] = x;
                xy[2 * pos + 1] = y;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) res[0] = s_base;
}

// predicted (distorted) pixel of feature i at state x, with the reference's two visibility gates (src/ExtendKF.cpp:56-135)
__device__ __forceinline__ bool predict_pixel_dev(const CamDev& cam, const double* x, int off, int type, double& hx, double& hy) {
    double R[9];
    q2r_dev(x + 3, R);
    double d[3];
    if (type == 0) {
        const double th = x[off + 3], ph = x[off + 4], rho = x[off + 5];
        const double mi[3] = {cos(ph) * sin(th), -sin(ph), cos(ph) * cos(th)};
        for (int k = 0; k < 3; k++) d[k] = (x[off + k] - x[k]) * rho + mi[k];
    } else {
        for (int k = 0; k < 3; k++) d[k] = x[off + k] - x[k];
    }
    double hrl[3];
    if (type == 0) {
        for (int k = 0; k < 3; k++) hrl[k] = R[k] * d[0] + R[3 + k] * d[1] + R[6 + k] * d[2];  // R^T d (:75)
    } else {
        double Ri[9];
        inv3_dev(R, Ri);
        for (int k = 0; k < 3; k++) hrl[k] = Ri[3 * k] * d[0] + Ri[3 * k + 1] * d[1] + Ri[3 * k + 2] * d[2];  // R^-1 d (:83)
    }
    const double ax = atan2(hrl[0], hrl[2]) * 180 / M_PI, ay = atan2(hrl[1], hrl[2]) * 180 / M_PI;
    if (ax < -60 || ax > 60 || ay < -60 || ay > 60) return false;
    const double uu = cam.Cx + (hrl[0] / hrl[2]) * cam.f * (1.0 / cam.dx);
    const double vu = cam.Cy + (hrl[1] / hrl[2]) * cam.f * (1.0 / cam.dy);
    distort_dev(cam, uu, vu, hx, hy);
    return (hx > 0) && (hx < cam.nCols) && (hy > 0) && (hy < cam.nRows);
}

// step 3 of Map::initialize_a_features (src/Map.cpp:248-261): how many features predicted at x_k_k fall inside the sampling box
__global__ void k_map_box_count(DevFilter* Fs, int b, CamDev cam, double cx, double cy, double sx, double sy, int* res) {
    const DevFilter& F = Fs[b];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F.N) return;
    double hx, hy;
    if (!predict_pixel_dev(cam, F.x_kk, F.foff[i], F.ftype[i], hx, hy)) return;
    if (hx > (cx - sx) && hx < (cx + sx) && hy > (cy - sy) && hy < (cy + sy)) atomicAdd(res, 1);
}

}  // namespace rslam
