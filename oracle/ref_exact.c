/*
 * CPU oracle, part B: plain-C restatement of the reference's image kernels on
 * the reference's own two layouts (packed RGBA8 as uint32, 1-channel fp64).
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into libmp_b200.so or the
 * `millipyde` extension.  Built by oracle/build_oracle.py into
 * oracle/_build/libref_exact.so and loaded with ctypes by tests/, smoke() and
 * bench.py's cpu_baseline leg.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference).  Arithmetic notes that matter for bit parity:
 *   - device code is compiled with FMA contraction on (nvcc and hipcc default),
 *     so `sum += a * b` in the reference kernels is one fused multiply-add;
 *     fma() is used explicitly here and this file must be compiled with
 *     -ffp-contract=off so nothing else fuses.
 *   - `powf`, `expf` take and return float; `255 * powf(..)` is a float
 *     product (int * float), only then widened to double.
 *   - float -> integer casts truncate toward zero.
 * Pinned by: tests/golden/ (outputs of the reference's own kernels, built
 * through the HIP->CUDA shim and run on a B200, see oracle/build_ref.py and
 * tests/golden/make_golden.py) and by the skimage restatement where the
 * reference's tests compare the two (decimal=4).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REF_RADIUS 8 /* src/include/millipyde_image.h:11 */
#define REF_TAPS (2 * REF_RADIUS + 1)

/* src/millipyde_image.cpp:50-66 -- luma of the first three bytes of each pixel,
 * divided by 255, capped at 1; alpha (if any) ignored. */
void ref_grey_u8(const uint8_t *rgb, double *grey, int width, int height, int channels)
{
    for (long i = 0; i < (long)width * height; ++i) {
        const uint8_t *p = rgb + i * channels;
        /* contraction as nvcc emits it for the reference's expression (pinned bit-for-bit by
         * tests/golden/reference_outputs.npz): g's product is rounded, r and b are fused */
        double acc = 0.7154 * p[1];
        acc = fma(0.2125, (double)p[0], acc);
        acc = fma(0.0721, (double)p[2], acc);
        grey[i] = fmin(1.0, acc / 255);
    }
}

/* src/millipyde_image.cpp:73-96 -- out[x][y] = in[y][x], element = elt bytes. */
void ref_transpose(const void *in, void *out, int width, int height, int elt)
{
    const char *s = (const char *)in;
    char *d = (char *)out;
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x)
            memcpy(d + ((long)x * height + y) * elt, s + ((long)y * width + x) * elt, elt);
}

/* src/millipyde_image.cpp:99-111 -- out[y][W-1-x] = in[y][x]. */
void ref_fliplr(const void *in, void *out, int width, int height, int elt)
{
    const char *s = (const char *)in;
    char *d = (char *)out;
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x)
            memcpy(d + ((long)y * width + (width - 1 - x)) * elt,
                   s + ((long)y * width + x) * elt, elt);
}

/* src/millipyde_image.cpp:114-139 with the degree->radian factor of :702.
 * Inverse-mapped nearest neighbour by int truncation about (W/2, H/2), zero
 * fill.  sin/cos here are glibc's; the device's may differ in the last ulp, so
 * pixels whose source coordinate lies within an ulp of an integer can differ
 * from a GPU run (tests allow a tiny mismatch fraction for this op only). */
void ref_rotate(const void *in, void *out, int width, int height, int elt, double angle_deg)
{
    const double angle = angle_deg * 0.01745329252;
    const double c = cos(angle), s = sin(angle);
    const double hw = (double)width / 2, hh = (double)height / 2;
    const char *src = (const char *)in;
    char *dst = (char *)out;
    for (int y = 0; y < height; ++y) {
        for (int x = 0; x < width; ++x) {
            /* same expression shape as the kernel; a*b - c*d + e contracts to
             * fma(a, b, -(c*d)) + e under nvcc */
            double t = ((double)y - hh) * s;
            int xr = (int)(fma((double)x - hw, c, -t) + hw);
            double u = ((double)y - hh) * c;
            int yr = (int)(fma((double)x - hw, s, u) + hh);
            char *o = dst + ((long)y * width + x) * elt;
            if (xr >= 0 && xr < width && yr >= 0 && yr < height)
                memcpy(o, src + ((long)yr * width + xr) * elt, elt);
            else
                memset(o, 0, elt);
        }
    }
}

/* src/millipyde_image.cpp:744-756 -- 17 weights, expf in float, normalised by
 * their double sum. */
void ref_gauss_weights(double sigma, double *w)
{
    double total = 0;
    for (int i = 0; i < REF_TAPS; ++i) {
        int dist = -1 * (REF_RADIUS - i);
        w[i] = expf((float)(-1 * ((dist * dist) / (2 * sigma * sigma))));
        total += w[i];
    }
    for (int i = 0; i < REF_TAPS; ++i)
        w[i] /= total;
}

/* src/millipyde_image.cpp:146-244 -- row pass into a scratch image, column pass
 * back into the input, zero padding, fused multiply-adds in tap order
 * k = -8..8 against w[8-k]; fmax(0, .) after the column pass only. */
void ref_gaussian_f64(double *img, int width, int height, double sigma)
{
    double w[REF_TAPS];
    ref_gauss_weights(sigma, w);
    double *tmp = (double *)malloc(sizeof(double) * (size_t)width * height);
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            double sum = 0;
            for (int k = -REF_RADIUS; k <= REF_RADIUS; ++k) {
                int xx = x + k;
                double v = (xx >= 0 && xx < width) ? img[(long)y * width + xx] : 0;
                sum = fma(v, w[REF_RADIUS - k], sum);
            }
            tmp[(long)y * width + x] = sum;
        }
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            double sum = 0;
            for (int k = -REF_RADIUS; k <= REF_RADIUS; ++k) {
                int yy = y + k;
                double v = (yy >= 0 && yy < height) ? tmp[(long)yy * width + x] : 0;
                sum = fma(v, w[REF_RADIUS - k], sum);
            }
            img[(long)y * width + x] = fmax(0.0, sum);
        }
    free(tmp);
}

/* src/millipyde_image.cpp:247-381 -- per byte lane: sum of (int)(byte * w)
 * (each product truncated before the integer add), & 0xff; the column pass
 * forces bits 24..31 to 0xff. */
void ref_gaussian_rgba(uint32_t *img, int width, int height, double sigma)
{
    double w[REF_TAPS];
    ref_gauss_weights(sigma, w);
    uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)width * height);
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            int lane[4] = {0, 0, 0, 0};
            for (int k = -REF_RADIUS; k <= REF_RADIUS; ++k) {
                int xx = x + k;
                uint32_t v = (xx >= 0 && xx < width) ? img[(long)y * width + xx] : 0;
                for (int b = 0; b < 4; ++b)
                    lane[b] += (int)(((v >> (8 * b)) & 0xff) * w[REF_RADIUS - k]);
            }
            tmp[(long)y * width + x] = ((uint32_t)(lane[3] & 0xff) << 24) |
                                       ((uint32_t)(lane[2] & 0xff) << 16) |
                                       ((uint32_t)(lane[1] & 0xff) << 8) |
                                       (uint32_t)(lane[0] & 0xff);
        }
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            int lane[3] = {0, 0, 0};
            for (int k = -REF_RADIUS; k <= REF_RADIUS; ++k) {
                int yy = y + k;
                uint32_t v = (yy >= 0 && yy < height) ? tmp[(long)yy * width + x] : 0;
                for (int b = 0; b < 3; ++b)
                    lane[b] += (int)(((v >> (8 * b)) & 0xff) * w[REF_RADIUS - k]);
            }
            img[(long)y * width + x] = 0xff000000u |
                                       ((uint32_t)(lane[2] & 0xff) << 16) |
                                       ((uint32_t)(lane[1] & 0xff) << 8) |
                                       (uint32_t)(lane[0] & 0xff);
        }
    free(tmp);
}

/* src/millipyde_image.cpp:384-398 */
void ref_brightness_f64(const double *in, double *out, long n, double delta)
{
    for (long i = 0; i < n; ++i) {
        double v = in[i] + delta;
        v = v < 0 ? 0 : v;
        v = v > 1 ? 1 : v;
        out[i] = v;
    }
}

/* src/millipyde_image.cpp:401-435 with delta8 = (char)(delta * 255) (:613).
 * R, G, B = byte lanes 0, 1, 2; alpha kept. */
void ref_brightness_rgba(const uint32_t *in, uint32_t *out, long n, double delta)
{
    const signed char d8 = (signed char)(delta * 255);
    for (long i = 0; i < n; ++i) {
        uint32_t px = in[i], res = px & 0xff000000u;
        for (int b = 0; b < 3; ++b) {
            int c = (int)((px >> (8 * b)) & 0xff) + d8;
            if (d8 > 0)
                c = c < 255 ? c : 255;
            else
                c = c > 0 ? c : 0;
            res |= (uint32_t)(c & 0xff) << (8 * b);
        }
        out[i] = res;
    }
}

/* src/millipyde_image.cpp:438-454 -- powf on float-converted operands. */
void ref_gamma_f64(const double *in, double *out, long n, double gamma, double gain)
{
    for (long i = 0; i < n; ++i) {
        double v = gain * powf((float)in[i], (float)gamma);
        v = v < 0 ? 0 : v;
        v = v > 1 ? 1 : v;
        out[i] = v;
    }
}

/* src/millipyde_image.cpp:457-491 -- gain * (255 * powf(c / 255., gamma)) where
 * 255 * powf() is a float product; clamp to [0, 255]; truncating cast. */
void ref_gamma_rgba(const uint32_t *in, uint32_t *out, long n, double gamma, double gain)
{
    uint8_t lut[256];
    for (int c = 0; c < 256; ++c) {
        float p = powf((float)((double)c / 255), (float)gamma);
        double t = gain * (double)(255.0f * p);
        t = t < 0 ? 0 : t;
        lut[c] = t > 255 ? 255 : (uint8_t)t;
    }
    for (long i = 0; i < n; ++i) {
        uint32_t px = in[i];
        out[i] = (px & 0xff000000u) | ((uint32_t)lut[(px >> 16) & 0xff] << 16) |
                 ((uint32_t)lut[(px >> 8) & 0xff] << 8) | lut[px & 0xff];
    }
}

/* src/millipyde_image.cpp:494-524 */
void ref_colorize_rgba(const uint32_t *in, uint32_t *out, long n, double rm, double gm, double bm)
{
    const double m[3] = {rm, gm, bm};
    for (long i = 0; i < n; ++i) {
        uint32_t px = in[i], res = px & 0xff000000u;
        for (int b = 0; b < 3; ++b) {
            double t = (double)((px >> (8 * b)) & 0xff) * m[b];
            uint32_t c = t > 255 ? 255 : (uint8_t)t;
            res |= c << (8 * b);
        }
        out[i] = res;
    }
}
