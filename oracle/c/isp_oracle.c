/*
 * Plain-C (OpenMP) restatement of the reference's packed12 -> RGB camera-ISP path.
 * TEST INFRASTRUCTURE ONLY: it is the timed multi-core CPU baseline of bench.py ("kind": "port") and is
 * itself checked against the numpy oracle (tests/test_oracle_c.py).  The product never links it.
 *
 * Follows, stage by stage with the reference's materialised intermediates
 * (paths relative to /root/reference/taichi_image):
 *   decode12 scaled          packed.py:23-31, :98-100, :108-117
 *   Malvar demosaic + CCM    bayer.py:30-55 (tables), :137-155 (filter_at), :158-175
 *   bilinear resize          interpolate.py:19-34, :59-66 via camera_isp.py:302-315 (before metering / tone map)
 *   metering + moving avg    camera_isp.py:102-175
 *   Reinhard                 camera_isp.py:177-218
 *   linear                   tonemap.py:11-17 via camera_isp.py:220-227
 * Build: gcc -O2 -fopenmp -ffp-contract=off (no FMA contraction, no fast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const signed char TAPS[4][13][3] = {
  {{0,-2,-3},{0,0,4},{0,4,0},{0,0,4},{0,-2,-3},{0,4,0},{16,8,12},{0,4,0},{0,-2,-3},{0,0,4},{0,4,0},{0,0,4},{0,-2,-3}},
  {{-2,0,1},{-2,0,-2},{8,0,0},{-2,0,-2},{1,0,-2},{0,0,8},{10,16,10},{0,0,8},{1,0,-2},{-2,0,-2},{8,0,0},{-2,0,-2},{-2,0,1}},
  {{1,0,-2},{-2,0,-2},{0,0,8},{-2,0,-2},{-2,0,1},{8,0,0},{10,16,10},{8,0,0},{-2,0,1},{-2,0,-2},{0,0,8},{-2,0,-2},{1,0,-2}},
  {{-3,-2,0},{4,0,0},{0,4,0},{4,0,0},{-3,-2,0},{0,4,0},{12,8,16},{0,4,0},{-3,-2,0},{4,0,0},{0,4,0},{4,0,0},{-3,-2,0}}};
static const signed char D0[13] = {-2,-1,-1,-1,0,0,0,0,0,1,1,1,2};
static const signed char D1[13] = {0,-1,0,1,-2,-1,0,1,2,-1,0,1,0};
/* kernel_patterns (bayer.py:92-97) indexed by BayerPattern value RGGB=0 GRBG=1 GBRG=2 BGGR=3, slot = (row&1)+2*(col&1) */
static const int KPAT[4][4] = {{0,1,2,3},{2,3,0,1},{1,0,3,2},{3,2,1,0}};

static inline float round_isp(float x, int cam16) { return cam16 ? (float)(_Float16)x : x; }
static inline float clamp01(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
static inline float gray_of(float r, float g, float b) { return (r * 0.299f + g * 0.587f) + b * 0.114f; }

static void decode_frame(const uint8_t* p, float* cfa, int H, int W, int cam16) {
  const float k = (float)(1.0 / 4095.0);
  const long pairs = (long)H * W / 2;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < pairs; ++i) {
    const unsigned b0 = p[3 * i], b1 = p[3 * i + 1], b2 = p[3 * i + 2];
    cfa[2 * i] = round_isp((float)(((b1 & 0xF) << 8) | b0) * k, cam16);
    cfa[2 * i + 1] = round_isp((float)((b2 << 4) | (b1 >> 4)) * k, cam16);
  }
}

static void demosaic_frame(const float* cfa, float* rgb, int H, int W, int pattern, const float* m, int cam16) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < H; ++r) {
    for (int c = 0; c < W; ++c) {
      const int K = KPAT[pattern][(r & 1) + 2 * (c & 1)];
      float acc[3] = {0.f, 0.f, 0.f}, t[3] = {0.f, 0.f, 0.f};
      for (int i = 0; i < 13; ++i) {
        const int rr = r + D0[i], cc = c + D1[i];
        if (rr >= 0 && rr < H && cc >= 0 && cc < W) {
          const float v = cfa[(long)rr * W + cc];
          for (int k = 0; k < 3; ++k) {
            const float w = (float)TAPS[K][i][k];
            acc[k] = acc[k] + v * w;
            t[k] += w;
          }
        }
      }
      float x = acc[0] / (1.0f * t[0]), y = acc[1] / (1.0f * t[1]), z = acc[2] / (1.0f * t[2]);
      if (m) {
        const float a = (x * m[0] + y * m[1]) + z * m[2];
        const float b = (x * m[3] + y * m[4]) + z * m[5];
        const float d = (x * m[6] + y * m[7]) + z * m[8];
        x = a; y = b; z = d;
      }
      float* o = rgb + ((long)r * W + c) * 3;
      o[0] = round_isp(clamp01(x), cam16);
      o[1] = round_isp(clamp01(y), cam16);
      o[2] = round_isp(clamp01(z), cam16);
    }
  }
}

/* interpolate.py:59-66: p = I / scale, p1 = trunc(p), clamp-to-edge taps, mix along dim 0 first, cast to the ISP dtype */
static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static void resize_frame(const float* src, float* dst, int H, int W, int Ho, int Wo, float scale_r, float scale_c, int cam16) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < Ho; ++r) {
    const float pr = (float)r / scale_r;
    const int r1 = (int)pr;
    const float fr = pr - (float)r1;
    const int ra = r1 < 0 ? 0 : (r1 > H - 1 ? H - 1 : r1), rb = r1 + 1 < 0 ? 0 : (r1 + 1 > H - 1 ? H - 1 : r1 + 1);
    for (int c = 0; c < Wo; ++c) {
      const float pc = (float)c / scale_c;
      const int c1 = (int)pc;
      const float fc = pc - (float)c1;
      const int ca = c1 < 0 ? 0 : (c1 > W - 1 ? W - 1 : c1), cb = c1 + 1 < 0 ? 0 : (c1 + 1 > W - 1 ? W - 1 : c1 + 1);
      for (int k = 0; k < 3; ++k) {
        const float y1 = mixf(src[((long)ra * W + ca) * 3 + k], src[((long)rb * W + ca) * 3 + k], fr);
        const float y2 = mixf(src[((long)ra * W + cb) * 3 + k], src[((long)rb * W + cb) * 3 + k], fr);
        dst[((long)r * Wo + c) * 3 + k] = round_isp(mixf(y1, y2, fc), cam16);
      }
    }
  }
}

static void metering(float* const* rgb, int n, int H, int W, int stride, float alpha, float* metrics) {
  const int hs = (H + stride - 1) / stride, ws = (W + stride - 1) / stride;
  float mn = INFINITY, mx = -INFINITY;
#pragma omp parallel for reduction(min:mn) reduction(max:mx) schedule(static)
  for (long q = 0; q < (long)n * hs; ++q) {
    const int f = (int)(q / hs), i = (int)(q % hs);
    for (int j = 0; j < ws; ++j) {
      const float* p = rgb[f] + ((long)i * stride * W + (long)j * stride) * 3;
      for (int k = 0; k < 3; ++k) { mn = fminf(mn, p[k]); mx = fmaxf(mx, p[k]); }
    }
  }
  const float bmin = mn + alpha * (metrics[0] - mn), bmax = mx + alpha * (metrics[1] - mx);
  const float den = (bmax - bmin) + 1e-6f;
  float lmin = INFINITY, lmax = -INFINITY;
  double sl = 0, sg = 0, sr = 0, sgr = 0, sb = 0;
#pragma omp parallel for reduction(min:lmin) reduction(max:lmax) reduction(+:sl,sg,sr,sgr,sb) schedule(static)
  for (long q = 0; q < (long)n * hs; ++q) {
    const int f = (int)(q / hs), i = (int)(q % hs);
    for (int j = 0; j < ws; ++j) {
      const float* p = rgb[f] + ((long)i * stride * W + (long)j * stride) * 3;
      const float r = (p[0] - bmin) / den, g = (p[1] - bmin) / den, b = (p[2] - bmin) / den;
      const float gray = gray_of(r, g, b);
      const float lg = logf(fmaxf(gray, 1e-4f));
      lmin = fminf(lmin, lg); lmax = fmaxf(lmax, lg);
      sl += lg; sg += gray; sr += r; sgr += g; sb += b;
    }
  }
  const float cnt = (float)((long)n * hs * ws);
  const float stats[9] = {bmin, bmax, lmin, lmax, (float)sl / cnt, (float)sg / cnt, (float)sr / cnt, (float)sgr / cnt, (float)sb / cnt};
  for (int k = 0; k < 9; ++k) metrics[k] = stats[k] + alpha * (metrics[k] - stats[k]);
}

static inline void store_q(void* out, long idx, float q, int out_u16) {
  q = (q > 0.f) ? q : 0.f;              /* negative / NaN -> 0 */
  if (q > 1.f) q = 1.f;
  if (out_u16) ((uint16_t*)out)[idx] = (uint16_t)(65535.f * q);
  else ((uint8_t*)out)[idx] = (uint8_t)(255.f * q);
}

static void tonemap_linear(const float* rgb, void* out, long n, const float* metrics, float gamma, int out_u16) {
  const float bmin = metrics[0], inv = 1.0f / (metrics[1] - metrics[0]), ig = 1.0f / gamma;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    float x = (rgb[i] - bmin) * inv;
    if (gamma != 1.0f) x = powf(x, ig);
    store_q(out, i, x, out_u16);
  }
}

static void tonemap_reinhard(float* rgb, void* out, long npx, const float* m, float gamma, float intensity, float la,
                             float ca, int cam16, int out_u16) {
  const float bmin = m[0], range = m[1] - m[0];
  const float key = (m[3] - m[4]) / (m[3] - m[2]);
  const float map_key = 0.3f + 0.7f * powf(key, 1.4f);
  const float ki = expf(-intensity);
  float mean[3];
  for (int k = 0; k < 3; ++k) mean[k] = m[5] + ca * (m[6 + k] - m[5]);
  float mx = 1e-6f;
#pragma omp parallel for reduction(max:mx) schedule(static)
  for (long i = 0; i < npx; ++i) {
    float* p = rgb + 3 * i;
    float s[3];
    for (int k = 0; k < 3; ++k) s[k] = (p[k] - bmin) / range;
    const float gray = gray_of(s[0], s[1], s[2]);
    for (int k = 0; k < 3; ++k) {
      const float ac = gray + ca * (s[k] - gray);
      const float am = mean[k] + la * (ac - mean[k]);
      const float adapt = powf(ki * am, map_key);
      const float v = s[k] * (1.0f / (adapt + s[k]));
      mx = fmaxf(mx, v);
      p[k] = round_isp(v, cam16);
    }
  }
  const float ig = (float)(1.0 / (double)gamma);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < npx * 3; ++i) {
    float q = rgb[i] / mx;
    if (gamma != 1.0f) q = powf(q, ig);
    store_q(out, i, q, out_u16);
  }
}

/* returns 0 on success.  metrics: 9 floats in/out; alpha = weight of the previous metrics. */
static int process_impl(const uint8_t* const* packed, void* const* out, int n_frames, int H, int W, int pattern,
                        int cam16, int out_u16, int reinhard, const float* ccm, float gamma, float intensity,
                        float la, float ca, int stride, float alpha, float* metrics, int nthreads,
                        int Ho, int Wo, float scale_r, float scale_c) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  float** rgb = (float**)malloc(sizeof(float*) * n_frames);
  float* cfa = (float*)malloc(sizeof(float) * (size_t)H * W);
  if (!rgb || !cfa) return -1;
  for (int f = 0; f < n_frames; ++f) {
    rgb[f] = (float*)malloc(sizeof(float) * (size_t)H * W * 3);
    if (!rgb[f]) return -1;
    decode_frame(packed[f], cfa, H, W, cam16);
    demosaic_frame(cfa, rgb[f], H, W, pattern, ccm, cam16);
    if (Ho > 0 && Wo > 0) {                 /* camera_isp.py:371-373: resize BEFORE metering / tone map */
      float* small = (float*)malloc(sizeof(float) * (size_t)Ho * Wo * 3);
      if (!small) return -1;
      resize_frame(rgb[f], small, H, W, Ho, Wo, scale_r, scale_c, cam16);
      free(rgb[f]);
      rgb[f] = small;
    }
  }
  if (Ho > 0 && Wo > 0) { H = Ho; W = Wo; }
  metering(rgb, n_frames, H, W, stride, alpha, metrics);
  for (int f = 0; f < n_frames; ++f) {
    if (reinhard) tonemap_reinhard(rgb[f], out[f], (long)H * W, metrics, gamma, intensity, la, ca, cam16, out_u16);
    else tonemap_linear(rgb[f], out[f], (long)H * W * 3, metrics, gamma, out_u16);
    free(rgb[f]);
  }
  free(rgb);
  free(cfa);
  return 0;
}

int isp_oracle_process(const uint8_t* const* packed, void* const* out, int n_frames, int H, int W, int pattern,
                       int cam16, int out_u16, int reinhard, const float* ccm, float gamma, float intensity,
                       float la, float ca, int stride, float alpha, float* metrics, int nthreads) {
  return process_impl(packed, out, n_frames, H, W, pattern, cam16, out_u16, reinhard, ccm, gamma, intensity, la, ca, stride, alpha,
                      metrics, nthreads, 0, 0, 1.f, 1.f);
}

/* the same with the ISP's bilinear resize to (Ho, Wo); out buffers are (Ho, Wo, 3) */
int isp_oracle_process_resized(const uint8_t* const* packed, void* const* out, int n_frames, int H, int W, int pattern,
                               int cam16, int out_u16, int reinhard, const float* ccm, float gamma, float intensity,
                               float la, float ca, int stride, float alpha, float* metrics, int nthreads,
                               int Ho, int Wo, float scale_r, float scale_c) {
  return process_impl(packed, out, n_frames, H, W, pattern, cam16, out_u16, reinhard, ccm, gamma, intensity, la, ca, stride, alpha,
                      metrics, nthreads, Ho, Wo, scale_r, scale_c);
}

int isp_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
