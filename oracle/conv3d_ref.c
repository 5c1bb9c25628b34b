/* ORACLE (test infrastructure only) -- plain-C direct Conv3d + BatchNorm3d(+ReLU), double accumulation.
 *
 * Independent restatement of what nn.Conv3d(k=(kt,kh,kw), padding=(0,p,p), stride 1) and
 * nn.BatchNorm3d (train: biased batch variance for normalisation, unbiased for the running update,
 * momentum 0.1, eps 1e-5) compute for the layers built at code/helpers/model.py:71-94 and applied at
 * code/helpers/model.py:111-149.  Used by tests/test_oracle.py to cross-check the torch functional calls in
 * oracle/slowfast_oracle.py on small shapes.  Layout NCDHW fp32 like the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

void conv3d_ncdhw_f32(const float *x, const float *w, const float *bias, float *y,
                      int64_t B, int64_t Cin, int64_t T, int64_t H, int64_t W,
                      int64_t Cout, int64_t kt, int64_t kh, int64_t kw, int64_t pad) {
    int64_t To = T - kt + 1;
    int64_t Ho = H + 2 * pad - kh + 1, Wo = W + 2 * pad - kw + 1;
    for (int64_t b = 0; b < B; ++b)
    for (int64_t co = 0; co < Cout; ++co)
    for (int64_t t = 0; t < To; ++t)
    for (int64_t h = 0; h < Ho; ++h)
    for (int64_t wq = 0; wq < Wo; ++wq) {
        double acc = bias ? (double)bias[co] : 0.0;
        for (int64_t ci = 0; ci < Cin; ++ci)
        for (int64_t a = 0; a < kt; ++a)
        for (int64_t i = 0; i < kh; ++i) {
            int64_t hi = h + i - pad;
            if (hi < 0 || hi >= H) continue;
            for (int64_t j = 0; j < kw; ++j) {
                int64_t wi = wq + j - pad;
                if (wi < 0 || wi >= W) continue;
                acc += (double)x[(((b * Cin + ci) * T + t + a) * H + hi) * W + wi] *
                       (double)w[(((co * Cin + ci) * kt + a) * kh + i) * kw + j];
            }
        }
        y[(((b * Cout + co) * To + t) * Ho + h) * Wo + wq] = (float)acc;
    }
}

/* y <- gamma*(y-mean)/sqrt(var+eps)+beta (optionally ReLU); train!=0 uses batch stats and updates running. */
void batchnorm3d_ncdhw_f32(float *y, const float *gamma, const float *beta, float *running_mean,
                           float *running_var, int64_t B, int64_t C, int64_t S, int train, int relu,
                           double momentum, double eps) {
    for (int64_t c = 0; c < C; ++c) {
        double mean, var;
        if (train) {
            double s = 0, ss = 0;
            for (int64_t b = 0; b < B; ++b) for (int64_t i = 0; i < S; ++i) s += y[(b * C + c) * S + i];
            mean = s / (double)(B * S);
            for (int64_t b = 0; b < B; ++b) for (int64_t i = 0; i < S; ++i) {
                double d = y[(b * C + c) * S + i] - mean; ss += d * d;
            }
            var = ss / (double)(B * S);
            double unbiased = (B * S > 1) ? ss / (double)(B * S - 1) : var;
            running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
            running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
        } else { mean = running_mean[c]; var = running_var[c]; }
        double inv = 1.0 / sqrt(var + eps);
        for (int64_t b = 0; b < B; ++b) for (int64_t i = 0; i < S; ++i) {
            double v = (y[(b * C + c) * S + i] - mean) * inv * gamma[c] + beta[c];
            if (relu && v < 0) v = 0;
            y[(b * C + c) * S + i] = (float)v;
        }
    }
}
