/* TEST INFRASTRUCTURE ONLY -- plain-C float64 restatement of the CTC alpha-beta recursion.
 *
 * The reference calls torch.nn.CTCLoss(reduction="sum") on log_softmax(logits)
 * (/root/reference/liteasr/criterions/hybrid_ctc_attn.py:32,67-75).  The arithmetic lives in a
 * third-party dependency (torch, unpinned by the reference; container pin torch 2.11.0: ATen
 * native/LossCTC.cpp, not vendored under /root/reference), so this file restates the published
 * algorithm (Graves et al. 2006, eqs. 6-8, 10-11, 16) in log space and is pinned against
 * torch.nn.CTCLoss outputs stored in tests/golden/ctc_golden.json (tests/test_oracle_golden.py).
 *
 * Contract (one utterance b at a time, OpenMP over b):
 *   lp      (T,B,V) log-probabilities, row-major
 *   targets (B,Lmax) int64, padding ignored beyond tgt_len[b]
 *   nll[b]  = -log p(l_b | x_b)   (+inf if no valid alignment)
 *   grad    (T,B,V) d nll[b] / d lp[t,b,c] = -exp(LSE_{s: l'_s = c}(alpha_t(s)+beta_t(s)) - lp[t,b,c] + nll[b])
 *           and 0 for t >= in_len[b]   (NaN rows if infeasible, like zero_infinity=False)
 * Product code never links or calls this file.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static double lse2(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    double m = a > b ? a : b;
    return m + log(exp(a - m) + exp(b - m));
}
static double lse3(double a, double b, double c) { return lse2(lse2(a, b), c); }

int ctc_alpha_beta_f64(const double *lp, const long *targets, const long *in_len, const long *tgt_len,
                       double *nll, double *grad, double *unused_ws, long T, long B, long V, long Lmax,
                       int blank) {
    (void)unused_ws;
    int rc = 0;
    memset(grad, 0, sizeof(double) * (size_t)T * B * V);
#pragma omp parallel for schedule(dynamic)
    for (long b = 0; b < B; ++b) {
        const long Tb = in_len[b], L = tgt_len[b], S = 2 * L + 1;
        if (Tb > T || L > Lmax || Tb < 0 || L < 0) { rc = -1; continue; }
        if (Tb == 0) { nll[b] = (L == 0) ? 0.0 : INFINITY; continue; }
        long *ext = (long *)malloc(sizeof(long) * S);
        double *al = (double *)malloc(sizeof(double) * Tb * S);
        double *be = (double *)malloc(sizeof(double) * Tb * S);
        for (long s = 0; s < S; ++s) ext[s] = (s & 1) ? targets[b * Lmax + s / 2] : blank;
        for (long i = 0; i < Tb * S; ++i) { al[i] = -INFINITY; be[i] = -INFINITY; }
#define LP(t, c) lp[((t) * B + b) * V + (c)]
        al[0] = LP(0, blank);
        if (S > 1) al[1] = LP(0, ext[1]);
        for (long t = 1; t < Tb; ++t)
            for (long s = 0; s < S; ++s) {
                double a0 = al[(t - 1) * S + s];
                double a1 = s >= 1 ? al[(t - 1) * S + s - 1] : -INFINITY;
                double a2 = (s >= 2 && ext[s] != blank && ext[s] != ext[s - 2]) ? al[(t - 1) * S + s - 2] : -INFINITY;
                al[t * S + s] = LP(t, ext[s]) + lse3(a0, a1, a2);
            }
        double tot = S > 1 ? lse2(al[(Tb - 1) * S + S - 1], al[(Tb - 1) * S + S - 2]) : al[(Tb - 1) * S];
        nll[b] = -tot;
        if (isinf(tot)) {
            for (long t = 0; t < Tb; ++t)
                for (long c = 0; c < V; ++c) grad[(t * B + b) * V + c] = NAN;
        } else {
            be[(Tb - 1) * S + S - 1] = LP(Tb - 1, ext[S - 1]);
            if (S > 1) be[(Tb - 1) * S + S - 2] = LP(Tb - 1, ext[S - 2]);
            for (long t = Tb - 2; t >= 0; --t)
                for (long s = 0; s < S; ++s) {
                    double b0 = be[(t + 1) * S + s];
                    double b1 = s + 1 < S ? be[(t + 1) * S + s + 1] : -INFINITY;
                    double b2 = (s + 2 < S && ext[s] != blank && ext[s] != ext[s + 2]) ? be[(t + 1) * S + s + 2] : -INFINITY;
                    be[t * S + s] = LP(t, ext[s]) + lse3(b0, b1, b2);
                }
            for (long t = 0; t < Tb; ++t)
                for (long s = 0; s < S; ++s) {
                    double ab = al[t * S + s] + be[t * S + s];
                    if (ab > -INFINITY) grad[(t * B + b) * V + ext[s]] -= exp(ab - LP(t, ext[s]) - tot);
                }
        }
#undef LP
        free(ext); free(al); free(be);
    }
    return rc;
}
