/*
 * rpt_oracle.c -- CPU restatement of the ocramz/rp-tree hot path.  TEST INFRASTRUCTURE ONLY
 * (see rpt_oracle.h for who may use it and for the "parity unpinned" statement).
 *
 * Deliberately mirrors the reference's *algorithm*: a recursive Bin/Tip tree, a full stable merge
 * sort of (point, projection) pairs at every node, right-fold sparse dot, left-fold distance sum.
 * It is not meant to be fast.  Build with -ffp-contract=off.
 */
#include "rpt_oracle.h"
#include <stdlib.h>
#include <pthread.h>
#include <string.h>
#include <math.h>

/* ------------------------------------------------------------------------------------------------
 * SplitMix64 -- Hackage splitmix (System.Random.SplitMix), pinned by stack.yaml:22 (LTS-17.10).
 * Known answer (haddock): mkSMGen 42 == SMGen 9297814886316923340 13679457532755275413.
 * ---------------------------------------------------------------------------------------------- */
static uint64_t shift_xor(int n, uint64_t w) { return w ^ (w >> n); }
static uint64_t shift_xor_mul(int n, uint64_t k, uint64_t w) { return shift_xor(n, w) * k; }

uint64_t orc_mix64(uint64_t z0) {
    uint64_t z1 = shift_xor_mul(33, 0xff51afd7ed558ccdULL, z0);
    uint64_t z2 = shift_xor_mul(33, 0xc4ceb9fe1a85ec53ULL, z1);
    return shift_xor(33, z2);
}
static uint64_t mix64variant13(uint64_t z0) {
    uint64_t z1 = shift_xor_mul(30, 0xbf58476d1ce4e5b9ULL, z0);
    uint64_t z2 = shift_xor_mul(27, 0x94d049bb133111ebULL, z1);
    return shift_xor(31, z2);
}
uint64_t orc_mix_gamma(uint64_t z0) {
    uint64_t z1 = mix64variant13(z0) | 1ULL;
    int n = __builtin_popcountll(z1 ^ (z1 >> 1));
    return n >= 24 ? z1 : z1 ^ 0xaaaaaaaaaaaaaaaaULL;
}
void orc_mk_smgen(uint64_t s, orc_smgen* g) {
    g->seed = orc_mix64(s);
    g->gamma = orc_mix_gamma(s + 0x9e3779b97f4a7c15ULL);
}
uint64_t orc_next_word64(orc_smgen* g) {
    g->seed += g->gamma;
    return orc_mix64(g->seed);
}
double orc_next_double(orc_smgen* g) {
    return (double)(orc_next_word64(g) >> 11) * 0x1.0p-53;
}

/* Data.Number.Erf (erf package) invnormcdf for Double, as recalled: Acklam's rational approximation
 * followed by one Halley step.  UNVERIFIED against the Hackage source (no network, not vendored). */
static double acklam_inorm(double p) {
    static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                                3.754408661907416e+00};
    const double plow = 0.02425, phigh = 1 - plow;
    if (p < plow) {
        double q = sqrt(-2 * log(p));
        return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
               ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    } else if (p <= phigh) {
        double q = p - 0.5, r = q * q;
        return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
               (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
    } else {
        double q = sqrt(-2 * log(1 - p));
        return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
               ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
}
double orc_invnormcdf(double p) {
    if (p == 0) return -INFINITY;
    if (p == 1) return INFINITY;
    double x = acklam_inorm(p);
    double e = 0.5 * erfc(-x / sqrt(2.0)) - p;
    double u = e * sqrt(2 * M_PI) * exp(x * x / 2);
    return x - u / (1 + x * u / 2);
}
/* splitmix-distributions-0.9.0.0 stdNormal = normal 0 1 (one uniform through the inverse CDF). */
double orc_std_normal(orc_smgen* g) {
    double u = orc_next_double(g);
    return orc_invnormcdf(u) * 1.0 + 0.0;
}

/* ------------------------------------------------------------------------------------------------
 * Vector algebra -- src/Data/RPTree/Internal.hs
 * ---------------------------------------------------------------------------------------------- */

/* innerSD, Internal.hs:369-382:  go i | i >= nz1 || i >= nz2 = 0 | otherwise = (xl * xr +) $ go (succ i)
 * i.e. x0*y0 + (x1*y1 + (... + (x_{z-1}*y_{z-1} + 0))): a RIGHT fold, evaluated innermost first. */
double orc_inner_sd(int64_t nz, const int32_t* idx, const double* val, const double* x, int64_t d) {
    int64_t z = nz < d ? nz : d;     /* the `i >= nz2` guard compares the POSITION i with the dense length */
    double acc = 0.0;
    for (int64_t i = z - 1; i >= 0; --i) {
        double prod = val[i] * x[idx[i]];
        acc = prod + acc;
    }
    return acc;
}

/* innerSD of one hyperplane against every row of X (n x d row-major): the keys partitionAtMedian sorts by
 * (Internal.hs:504).  Used by the full-size property tests. */
void orc_project_all(int64_t nz, const int32_t* idx, const double* val, const double* X, int64_t n, int64_t d, double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = orc_inner_sd(nz, idx, val, X + i * d, d);
}

/* innerSS, Internal.hs:351-366: merge join, right fold over the matches. */
double orc_inner_ss(int64_t nz1, const int32_t* i1, const double* v1, int64_t nz2, const int32_t* i2, const double* v2) {
    /* collect matches left to right, then fold from the right */
    int64_t cap = nz1 < nz2 ? nz1 : nz2, m = 0;
    double* prods = (double*)malloc(sizeof(double) * (size_t)(cap > 0 ? cap : 1));
    int64_t a = 0, b = 0;
    while (a < nz1 && b < nz2) {
        if (i1[a] == i2[b]) { prods[m++] = v1[a] * v2[b]; ++a; ++b; }
        else if (i1[a] < i2[b]) ++a;
        else ++b;
    }
    double acc = 0.0;
    for (int64_t i = m - 1; i >= 0; --i) acc = prods[i] + acc;
    free(prods);
    return acc;
}

/* innerDD, Internal.hs:384-385: VG.sum (zipWith (*)) -- a strict LEFT fold from 0. */
double orc_inner_dd(const double* u, const double* v, int64_t d) {
    double acc = 0.0;
    for (int64_t i = 0; i < d; ++i) { double p = u[i] * v[i]; acc = acc + p; }
    return acc;
}

static int g_use_pow = 0;
void orc_set_use_pow(int on) { g_use_pow = on; }

/* metricDDL2, Internal.hs:403-406: sqrt $ VG.sum $ VG.map (** 2) (zipWith (-) u v).
 * `x ** 2` is libm pow(x,2) in GHC; x*x is the correctly rounded square that pow approximates to <1ulp.
 * Default: x*x (deterministic, what the CUDA path computes).  orc_set_use_pow(1) switches to pow. */
double orc_metric_dd_l2(const double* u, const double* v, int64_t d) {
    double acc = 0.0;
    for (int64_t i = 0; i < d; ++i) {
        double df = u[i] - v[i];
        double sq = g_use_pow ? pow(df, 2.0) : df * df;
        acc = acc + sq;
    }
    return sqrt(acc);
}

/* metricSSL2, Internal.hs:389-393: sqrt $ VG.sum $ VG.map (\(_, x) -> x ** 2) (u `diffSS` v) with
 * diffSS = binSS (-) 0 (Internal.hs:432-453): a merge join that STOPS when either operand is exhausted, so the
 * components of the longer-lasting vector beyond the other's last index are dropped (the reference's quirk). */
double orc_metric_ss_l2(int64_t nz1, const int32_t* i1, const double* v1, int64_t nz2, const int32_t* i2, const double* v2) {
    int64_t a = 0, b = 0;
    double acc = 0.0;
    while (a < nz1 && b < nz2) {
        double df;
        if (i1[a] == i2[b])     { df = v1[a] - v2[b]; ++a; ++b; }
        else if (i1[a] < i2[b]) { df = v1[a] - 0.0; ++a; }
        else                    { df = 0.0 - v2[b]; ++b; }
        acc = acc + (g_use_pow ? pow(df, 2.0) : df * df);
    }
    return sqrt(acc);
}

/* binSDD, Internal.hs:455-470 -- note the quirk: stops when EITHER operand is exhausted. */
static int64_t bin_sdd(int sub, int64_t nz, const int32_t* idx, const double* val, const double* x, int64_t d, double* out) {
    int64_t i1 = 0, i2 = 0, m = 0;
    while (i1 < nz && i2 < d) {
        int64_t il = idx[i1];
        double xl = val[i1], xr = x[i2];
        if (il == i2)      { out[m++] = sub ? xl - xr : xl + xr; ++i1; ++i2; }
        else if (il < i2)  { out[m++] = sub ? xl - 0.0 : xl + 0.0; ++i1; }
        else               { out[m++] = sub ? 0.0 - xr : 0.0 + xr; ++i2; }
    }
    return m;
}
int64_t orc_sum_sd (int64_t nz, const int32_t* idx, const double* val, const double* x, int64_t d, double* out) { return bin_sdd(0, nz, idx, val, x, d, out); }
int64_t orc_diff_sd(int64_t nz, const int32_t* idx, const double* val, const double* x, int64_t d, double* out) { return bin_sdd(1, nz, idx, val, x, d, out); }
/* metricSDL2, Internal.hs:396-400: sqrt $ VG.sum $ VG.map (** 2) (u `diffSD` v): the dense operand is only
 * visited up to the sparse operand's last index (binSDD stops when the sparse side is exhausted). */
double orc_metric_sd_l2(int64_t nz, const int32_t* idx, const double* val, const double* x, int64_t d) {
    double* tmp = (double*)malloc(sizeof(double) * (size_t)(nz + d + 1));
    int64_t m = bin_sdd(1, nz, idx, val, x, d, tmp);
    double acc = 0.0;
    for (int64_t i = 0; i < m; ++i) acc = acc + (g_use_pow ? pow(tmp[i], 2.0) : tmp[i] * tmp[i]);
    free(tmp);
    return sqrt(acc);
}

/* rpTreeCfg, Conduit.hs:132-141 */
void orc_rptree_cfg(int64_t minl, int64_t n, int64_t d, int64_t* maxd, int64_t* nchunk, double* pnz) {
    /* logBase b x = log x / log b in GHC's Floating Double instance */
    *maxd = (int64_t)ceil(log((double)n / (double)minl) / log(2.0));
    *nchunk = (int64_t)ceil((double)n / 100.0);
    double pmin = 1.0 / (log((double)d) / log(10.0));
    *pnz = pmin < 1.0 ? pmin : 1.0;
}

/* ------------------------------------------------------------------------------------------------
 * Hyperplanes -- sparseVG (Gen.hs:178-195) replicated maxd times per tree, T trees, one generator
 * (Batch.hs:57-63): for i in [0,dim): flag <- bernoulli p; if flag then x <- stdNormal.
 * ---------------------------------------------------------------------------------------------- */
int64_t orc_gen_hyperplanes(uint64_t seed, int32_t T, int32_t maxd, double pnz, int32_t dim,
                            int64_t* off, int32_t* idx, double* val) {
    orc_smgen g; orc_mk_smgen(seed, &g);
    int64_t nnz = 0;
    for (int32_t t = 0; t < T; ++t)
        for (int32_t l = 0; l < maxd; ++l) {
            if (off) off[(int64_t)t * maxd + l] = nnz;
            for (int32_t i = 0; i < dim; ++i) {
                double u = orc_next_double(&g);
                if (u < pnz) {
                    double x = orc_std_normal(&g);
                    if (idx) { idx[nnz] = i; val[nnz] = x; }
                    ++nnz;
                }
            }
        }
    if (off) off[(int64_t)T * maxd] = nnz;
    return nnz;
}

/* ------------------------------------------------------------------------------------------------
 * Forest
 * ---------------------------------------------------------------------------------------------- */
typedef struct onode {
    int is_bin;
    double thr, mlo, mhi;
    struct onode *l, *r;
    uint32_t* ids; int64_t n;       /* Tip payload (row ids standing in for Embed values) */
} onode;

struct orc_forest {
    const double* X; int64_t n; int32_t d; int32_t T, maxd, minl;
    const int64_t* sp_off; const int32_t* sp_idx; const double* sp_val;   /* data points as SVectors (CSR) when X == NULL */
    int64_t* hp_off; int32_t* hp_idx; double* hp_val;
    onode** roots;
};

static onode* tip_new(const uint32_t* ids, int64_t n) {
    onode* t = (onode*)calloc(1, sizeof(onode));
    t->n = n;
    t->ids = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    if (n > 0) memcpy(t->ids, ids, sizeof(uint32_t) * (size_t)n);
    return t;
}
static void node_free(onode* t) {
    if (!t) return;
    if (t->is_bin) { node_free(t->l); node_free(t->r); }
    free(t->ids); free(t);
}

typedef struct { uint32_t id; double key; } pk;

/* a query point: DVector (dense != NULL) or SVector (nz, idx, val) */
typedef struct { const double* dense; int64_t nz; const int32_t* idx; const double* val; } qref;

/* r `inner` x for a data point: innerSD (dense data) or innerSS (sparse data) -- instances Internal.hs:322-341 */
static double point_proj(const orc_forest* f, int64_t hp, uint32_t id) {
    int64_t s = f->hp_off[hp], e = f->hp_off[hp + 1];
    if (f->X) return orc_inner_sd(e - s, f->hp_idx + s, f->hp_val + s, f->X + (int64_t)id * f->d, f->d);
    int64_t a = f->sp_off[id], b = f->sp_off[id + 1];
    return orc_inner_ss(e - s, f->hp_idx + s, f->hp_val + s, b - a, f->sp_idx + a, f->sp_val + a);
}
static double query_proj(const orc_forest* f, int64_t hp, const qref* q) {
    int64_t s = f->hp_off[hp], e = f->hp_off[hp + 1];
    if (q->dense) return orc_inner_sd(e - s, f->hp_idx + s, f->hp_val + s, q->dense, f->d);
    return orc_inner_ss(e - s, f->hp_idx + s, f->hp_val + s, q->nz, q->idx, q->val);
}
/* eEmbed xe `metricL2` q: metricDDL2 / metricSDL2 / metricSSL2 by the operand types (Internal.hs:322-341) */
static double point_dist(const orc_forest* f, uint32_t id, const qref* q) {
    if (f->X) return orc_metric_dd_l2(f->X + (int64_t)id * f->d, q->dense, f->d);
    int64_t a = f->sp_off[id], b = f->sp_off[id + 1];
    if (q->dense) return orc_metric_sd_l2(b - a, f->sp_idx + a, f->sp_val + a, q->dense, f->d);
    return orc_metric_ss_l2(b - a, f->sp_idx + a, f->sp_val + a, q->nz, q->idx, q->val);
}

/* Ord Double compare as GHC defines it: LT if a<b, EQ if a==b, else GT (so NaN -> GT). */
static int cmp_double(double a, double b) { return a < b ? -1 : (a == b ? 0 : 1); }

/* stable top-down merge sort by key (stands in for Data.Vector.Algorithms.Merge.sortBy, which is stable) */
static void msort(pk* a, pk* tmp, int64_t n) {
    if (n < 2) return;
    int64_t h = n / 2;
    msort(a, tmp, h); msort(a + h, tmp, n - h);
    int64_t i = 0, j = h, k = 0;
    while (i < h && j < n) {
        if (cmp_double(a[j].key, a[i].key) < 0) tmp[k++] = a[j++];   /* take right only if strictly smaller */
        else tmp[k++] = a[i++];
    }
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(pk) * (size_t)n);
}

/* partitionAtMedian, Internal.hs:484-505.  Returns 0 for Nothing. ids is replaced by the sorted order. */
static int partition_at_median(const orc_forest* f, int64_t hp, uint32_t* ids, int64_t n,
                               double* thr, double* mlo, double* mhi, int64_t* nh_out) {
    if (n < 1) return 0;
    pk* a = (pk*)malloc(sizeof(pk) * (size_t)n);
    pk* tmp = (pk*)malloc(sizeof(pk) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        a[i].id = ids[i];
        a[i].key = point_proj(f, hp, ids[i]);
    }
    msort(a, tmp, n);
    int64_t nh = n / 2;
    if (n >= 3)      { *mlo = a[nh - 1].key; *mhi = a[nh + 1].key; }
    else if (n == 2) { *mlo = a[0].key;      *mhi = a[1].key; }
    else             { *mlo = a[0].key;      *mhi = a[0].key; }
    *thr = a[nh].key;
    *nh_out = nh;
    for (int64_t i = 0; i < n; ++i) ids[i] = a[i].id;
    free(a); free(tmp);
    return 1;
}

static double dmax(double x, double y) { return x <= y ? y : x; }   /* Haskell max */
static double dmin(double x, double y) { return x <= y ? x : y; }   /* Haskell min */

/* insert, Internal.hs:257-297.  Takes ownership of tt; xs (n row ids) is borrowed. */
static onode* insert_loop(const orc_forest* f, int32_t tree, int32_t lev, onode* tt, const uint32_t* xs, int64_t n) {
    int64_t hp = (int64_t)tree * f->maxd + lev;   /* r = rvs ! ixLev (lazy: only forced when used) */
    if (tt->is_bin) {
        if (lev >= f->maxd) return tt;
        uint32_t* w = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
        if (n > 0) memcpy(w, xs, sizeof(uint32_t) * (size_t)n);
        double thr, mlo, mhi; int64_t nh;
        if (!partition_at_median(f, hp, w, n, &thr, &mlo, &mhi, &nh)) {
            free(w); node_free(tt);
            return tip_new(NULL, 0);                 /* Nothing -> Tip () mempty  (Internal.hs:279) */
        }
        tt->mlo = dmax(tt->mlo, mlo);                /* margin0 <> margin : Max on lows ...        */
        tt->mhi = dmin(tt->mhi, mhi);                /* ... Min on highs (Internal.hs:86-87)        */
        tt->thr = (tt->thr + thr) / 2;               /* thr' = (thr0 + thr) / 2                     */
        tt->l = insert_loop(f, tree, lev + 1, tt->l, w, nh);
        tt->r = insert_loop(f, tree, lev + 1, tt->r, w + nh, n - nh);
        free(w);
        return tt;
    }
    /* Tip _ xs0 -> xs' = xs <> xs0 */
    int64_t n2 = n + tt->n;
    uint32_t* w = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n2 > 0 ? n2 : 1));
    if (n > 0) memcpy(w, xs, sizeof(uint32_t) * (size_t)n);
    if (tt->n > 0) memcpy(w + n, tt->ids, sizeof(uint32_t) * (size_t)tt->n);
    node_free(tt);
    if (lev >= f->maxd || n2 <= f->minl) {
        onode* t = tip_new(w, n2); free(w); return t;
    }
    double thr, mlo, mhi; int64_t nh;
    if (!partition_at_median(f, hp, w, n2, &thr, &mlo, &mhi, &nh)) { free(w); return tip_new(NULL, 0); }
    onode* b = (onode*)calloc(1, sizeof(onode));
    b->is_bin = 1; b->thr = thr; b->mlo = mlo; b->mhi = mhi;
    b->l = insert_loop(f, tree, lev + 1, tip_new(NULL, 0), w, nh);
    b->r = insert_loop(f, tree, lev + 1, tip_new(NULL, 0), w + nh, n2 - nh);
    free(w);
    return b;
}

orc_forest* orc_forest_new_chunked(const double* X, int64_t n, int32_t d, int32_t T, int32_t maxd, int32_t minl,
                                   int64_t chunk,
                                   const int64_t* hp_off, const int32_t* hp_idx, const double* hp_val) {
    orc_forest* f = (orc_forest*)calloc(1, sizeof(orc_forest));
    f->X = X; f->n = n; f->d = d; f->T = T; f->maxd = maxd; f->minl = minl;
    int64_t nhp = (int64_t)T * maxd, nnz = hp_off[nhp];
    f->hp_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nhp + 1));
    memcpy(f->hp_off, hp_off, sizeof(int64_t) * (size_t)(nhp + 1));
    f->hp_idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    f->hp_val = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    if (nnz > 0) { memcpy(f->hp_idx, hp_idx, sizeof(int32_t) * (size_t)nnz); memcpy(f->hp_val, hp_val, sizeof(double) * (size_t)nnz); }
    f->roots = (onode**)calloc((size_t)T, sizeof(onode*));
    if (chunk < 1) chunk = n > 0 ? n : 1;
    uint32_t* ids = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) ids[i] = (uint32_t)i;
    /* insertMulti maps over the IntMap of trees (Internal.hs:252-255); trees are independent, so looping
     * tree-outer / chunk-inner gives the same result as the reference's chunk-outer / tree-inner fold. */
    for (int32_t t = 0; t < T; ++t) {
        onode* tt = tip_new(NULL, 0);
        for (int64_t s = 0; s < n; s += chunk) {
            int64_t m = n - s < chunk ? n - s : chunk;
            tt = insert_loop(f, t, 0, tt, ids + s, m);
        }
        f->roots[t] = tt;
    }
    free(ids);
    return f;
}
/* forest over SVector data points (Embed SVector Double x): CSR rows with ascending indices; borrowed pointers */
orc_forest* orc_forest_new_sparse(int64_t n, int32_t d, const int64_t* sp_off, const int32_t* sp_idx, const double* sp_val,
                                  int32_t T, int32_t maxd, int32_t minl, int64_t chunk,
                                  const int64_t* hp_off, const int32_t* hp_idx, const double* hp_val) {
    /* build through the common path: X == NULL switches point_proj / point_dist to the sparse instances */
    orc_forest* f = (orc_forest*)calloc(1, sizeof(orc_forest));
    f->X = NULL; f->sp_off = sp_off; f->sp_idx = sp_idx; f->sp_val = sp_val;
    f->n = n; f->d = d; f->T = T; f->maxd = maxd; f->minl = minl;
    int64_t nhp = (int64_t)T * maxd, nnz = hp_off[nhp];
    f->hp_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nhp + 1));
    memcpy(f->hp_off, hp_off, sizeof(int64_t) * (size_t)(nhp + 1));
    f->hp_idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    f->hp_val = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    if (nnz > 0) { memcpy(f->hp_idx, hp_idx, sizeof(int32_t) * (size_t)nnz); memcpy(f->hp_val, hp_val, sizeof(double) * (size_t)nnz); }
    f->roots = (onode**)calloc((size_t)T, sizeof(onode*));
    if (chunk < 1) chunk = n > 0 ? n : 1;
    uint32_t* ids = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) ids[i] = (uint32_t)i;
    for (int32_t t = 0; t < T; ++t) {
        onode* tt = tip_new(NULL, 0);
        for (int64_t s = 0; s < n; s += chunk) {
            int64_t m = n - s < chunk ? n - s : chunk;
            tt = insert_loop(f, t, 0, tt, ids + s, m);
        }
        f->roots[t] = tt;
    }
    free(ids);
    return f;
}
/* The same forest with the independent trees built on `nthreads` host threads (createMulti is a map over the IntMap of
 * trees, Internal.hs:234-240; the reference itself is single threaded -- this exists so that the CPU arm of the benchmark
 * can use every host core).  Result identical to orc_forest_new. */
typedef struct { orc_forest* f; const uint32_t* ids; int32_t t0, t1, step; } mt_job;
static void* mt_run(void* arg) {
    mt_job* j = (mt_job*)arg;
    for (int32_t t = j->t0; t < j->t1; t += j->step)
        j->f->roots[t] = insert_loop(j->f, t, 0, tip_new(NULL, 0), j->ids, j->f->n);
    return NULL;
}
orc_forest* orc_forest_new_mt(const double* X, int64_t n, int32_t d, int32_t T, int32_t maxd, int32_t minl,
                              const int64_t* hp_off, const int32_t* hp_idx, const double* hp_val, int32_t nthreads) {
    orc_forest* f = (orc_forest*)calloc(1, sizeof(orc_forest));
    f->X = X; f->n = n; f->d = d; f->T = T; f->maxd = maxd; f->minl = minl;
    int64_t nhp = (int64_t)T * maxd, nnz = hp_off[nhp];
    f->hp_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nhp + 1));
    memcpy(f->hp_off, hp_off, sizeof(int64_t) * (size_t)(nhp + 1));
    f->hp_idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    f->hp_val = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    if (nnz > 0) { memcpy(f->hp_idx, hp_idx, sizeof(int32_t) * (size_t)nnz); memcpy(f->hp_val, hp_val, sizeof(double) * (size_t)nnz); }
    f->roots = (onode**)calloc((size_t)T, sizeof(onode*));
    uint32_t* ids = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) ids[i] = (uint32_t)i;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > T) nthreads = T;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    mt_job* jobs = (mt_job*)malloc(sizeof(mt_job) * (size_t)nthreads);
    for (int32_t i = 0; i < nthreads; ++i) {
        jobs[i].f = f; jobs[i].ids = ids; jobs[i].t0 = i; jobs[i].t1 = T; jobs[i].step = nthreads;
        pthread_create(&th[i], NULL, mt_run, &jobs[i]);
    }
    for (int32_t i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(th); free(jobs); free(ids);
    return f;
}
orc_forest* orc_forest_new(const double* X, int64_t n, int32_t d, int32_t T, int32_t maxd, int32_t minl,
                           const int64_t* hp_off, const int32_t* hp_idx, const double* hp_val) {
    return orc_forest_new_chunked(X, n, d, T, maxd, minl, n > 0 ? n : 1, hp_off, hp_idx, hp_val);
}
void orc_forest_free(orc_forest* f) {
    if (!f) return;
    for (int32_t t = 0; t < f->T; ++t) node_free(f->roots[t]);
    free(f->roots); free(f->hp_off); free(f->hp_idx); free(f->hp_val); free(f);
}

static int64_t count_points(const onode* t) { return t->is_bin ? count_points(t->l) + count_points(t->r) : t->n; }
static int64_t count_nodes(const onode* t) { return t->is_bin ? 1 + count_nodes(t->l) + count_nodes(t->r) : 1; }
int64_t orc_tree_size(const orc_forest* f, int32_t t) { return count_points(f->roots[t]); }

int64_t orc_tree_export(const orc_forest* f, int32_t t, int64_t* child, int32_t* depth, double* thr, double* mlo, double* mhi,
                        int64_t* seg_start, int64_t* seg_size, uint32_t* perm) {
    int64_t nn = count_nodes(f->roots[t]);
    if (!child) return nn;
    const onode** q = (const onode**)malloc(sizeof(onode*) * (size_t)nn);
    int64_t head = 0, tail = 0;
    q[tail] = f->roots[t]; depth[tail] = 0; seg_start[tail] = 0; ++tail;
    while (head < tail) {
        const onode* nd = q[head];
        int64_t g = head++;
        seg_size[g] = count_points(nd);
        if (nd->is_bin) {
            child[g] = tail;
            thr[g] = nd->thr; mlo[g] = nd->mlo; mhi[g] = nd->mhi;
            q[tail] = nd->l; depth[tail] = depth[g] + 1; seg_start[tail] = seg_start[g]; ++tail;
            q[tail] = nd->r; depth[tail] = depth[g] + 1; seg_start[tail] = seg_start[g] + count_points(nd->l); ++tail;
        } else {
            child[g] = -1; thr[g] = 0; mlo[g] = 0; mhi[g] = 0;
            if (perm && nd->n > 0) memcpy(perm + seg_start[g], nd->ids, sizeof(uint32_t) * (size_t)nd->n);
        }
    }
    free(q);
    return nn;
}

/* candidates, RPTree.hs:297-314 */
typedef struct { uint32_t* ids; int64_t n, cap; int count_only; } idbuf;
static void idbuf_push(idbuf* b, const uint32_t* ids, int64_t n) {
    if (!b->count_only) {
        if (b->n + n > b->cap) { b->cap = (b->n + n) * 2 + 16; b->ids = (uint32_t*)realloc(b->ids, sizeof(uint32_t) * (size_t)b->cap); }
        if (n > 0) memcpy(b->ids + b->n, ids, sizeof(uint32_t) * (size_t)n);
    }
    b->n += n;
}
static void cand_go(const orc_forest* f, int32_t tree, int32_t lev, const onode* tt, const qref* x, idbuf* out) {
    if (!tt->is_bin) { idbuf_push(out, tt->ids, tt->n); return; }
    int64_t hp = (int64_t)tree * f->maxd + lev;
    double proj = query_proj(f, hp, x);
    double dl = fabs(tt->mlo - proj), dr = fabs(tt->mhi - proj);
    if (proj < tt->thr && dl > dr)      { cand_go(f, tree, lev + 1, tt->l, x, out); cand_go(f, tree, lev + 1, tt->r, x, out); }
    else if (proj < tt->thr)            { cand_go(f, tree, lev + 1, tt->l, x, out); }
    else if (proj > tt->thr && dl < dr) { cand_go(f, tree, lev + 1, tt->l, x, out); cand_go(f, tree, lev + 1, tt->r, x, out); }
    else                                { cand_go(f, tree, lev + 1, tt->r, x, out); }
}
static int64_t candidates_q(const orc_forest* f, int32_t t, const qref* q, uint32_t* ids, int64_t cap) {
    idbuf b = {0};
    b.count_only = (ids == NULL);
    cand_go(f, t, 0, f->roots[t], q, &b);
    if (ids) { int64_t m = b.n < cap ? b.n : cap; if (m > 0) memcpy(ids, b.ids, sizeof(uint32_t) * (size_t)m); free(b.ids); }
    return b.n;
}

/* knn, RPTree.hs:174-176: all candidates of all trees (ascending tree key, duplicates kept),
 * stable sort by distance, take k.  dedup=1: knnPQ/nub semantics -- one entry per distinct DISTANCE
 * (heaps' Entry compares on priority only, RPTree.hs:187-194,224-227); the survivor of a group is
 * unpinned in the reference (heap internals), here: first in candidate order. */
int64_t orc_candidates(const orc_forest* f, int32_t t, const double* q, uint32_t* ids, int64_t cap) {
    qref r = {q, 0, NULL, NULL};
    return candidates_q(f, t, &r, ids, cap);
}
int64_t orc_candidates_sq(const orc_forest* f, int32_t t, int64_t qnz, const int32_t* qidx, const double* qval, uint32_t* ids, int64_t cap) {
    qref r = {NULL, qnz, qidx, qval};
    return candidates_q(f, t, &r, ids, cap);
}

static int64_t knn_q(const orc_forest* f, const qref* q, int32_t k, int32_t dedup, double* dist, uint32_t* ids) {
    idbuf b = {0};
    for (int32_t t = 0; t < f->T; ++t) cand_go(f, t, 0, f->roots[t], q, &b);
    int64_t c = b.n;
    pk* a = (pk*)malloc(sizeof(pk) * (size_t)(c > 0 ? c : 1));
    pk* tmp = (pk*)malloc(sizeof(pk) * (size_t)(c > 0 ? c : 1));
    for (int64_t i = 0; i < c; ++i) {
        a[i].id = b.ids[i];
        a[i].key = point_dist(f, b.ids[i], q);
    }
    msort(a, tmp, c);
    int64_t m = 0;
    for (int64_t i = 0; i < c && m < k; ++i) {
        if (dedup && i > 0 && cmp_double(a[i].key, a[i - 1].key) == 0) continue;
        dist[m] = a[i].key; ids[m] = a[i].id; ++m;
    }
    free(a); free(tmp); free(b.ids);
    return m;
}

int64_t orc_knn(const orc_forest* f, const double* q, int32_t k, int32_t dedup, double* dist, uint32_t* ids) {
    qref r = {q, 0, NULL, NULL};
    return knn_q(f, &r, k, dedup, dist, ids);
}
int64_t orc_knn_sq(const orc_forest* f, int64_t qnz, const int32_t* qidx, const double* qval, int32_t k, int32_t dedup, double* dist, uint32_t* ids) {
    qref r = {NULL, qnz, qidx, qval};
    return knn_q(f, &r, k, dedup, dist, ids);
}

/* candidatesH, RPTree.hs:318-341: every reached Tip with its priority p (p0 = 1/0; left: p `min` dl, right: p `min` dr) */
typedef struct { double p; const onode* tip; } hent;
typedef struct { hent* e; int64_t n, cap; } hbuf;
static void hbuf_push(hbuf* b, double p, const onode* tip) {
    if (b->n == b->cap) { b->cap = b->cap * 2 + 16; b->e = (hent*)realloc(b->e, sizeof(hent) * (size_t)b->cap); }
    b->e[b->n].p = p; b->e[b->n].tip = tip; ++b->n;
}
static void cand_h_go(const orc_forest* f, int32_t tree, int32_t lev, const onode* tt, const qref* x, double p, hbuf* out) {
    if (!tt->is_bin) { hbuf_push(out, p, tt); return; }
    int64_t hp = (int64_t)tree * f->maxd + lev;
    double proj = query_proj(f, hp, x);
    double dl = fabs(tt->mlo - proj), dr = fabs(tt->mhi - proj);
    double pl = dmin(p, dl), pr = dmin(p, dr);
    if (proj < tt->thr && dl > dr)      { cand_h_go(f, tree, lev + 1, tt->l, x, pl, out); cand_h_go(f, tree, lev + 1, tt->r, x, pr, out); }
    else if (proj < tt->thr)            { cand_h_go(f, tree, lev + 1, tt->l, x, pl, out); }
    else if (proj > tt->thr && dl < dr) { cand_h_go(f, tree, lev + 1, tt->l, x, pl, out); cand_h_go(f, tree, lev + 1, tt->r, x, pr, out); }
    else                                { cand_h_go(f, tree, lev + 1, tt->r, x, pr, out); }
}
/* knnH, RPTree.hs:199-217: pop the union heap in increasing priority, prepend each popped leaf, stop at the first pop
 * that would push the total past k once something has been taken.  The pop order among EQUAL priorities depends on the
 * internals of the `heaps` package (skew-binomial heap; not in /root/reference): UNPINNED.  This restatement breaks such
 * ties by (tree, leaf position left to right) -- the product documents and implements the same rule. */
static int64_t knn_h_q(const orc_forest* f, const qref* q, int32_t k, double* dist, uint32_t* ids, int64_t cap) {
    hbuf b = {0};
    for (int32_t t = 0; t < f->T; ++t) cand_h_go(f, t, 0, f->roots[t], q, 1.0 / 0.0, &b);
    /* stable insertion sort by priority (few entries) */
    for (int64_t i = 1; i < b.n; ++i) {
        hent v = b.e[i]; int64_t j = i;
        while (j > 0 && cmp_double(v.p, b.e[j - 1].p) < 0) { b.e[j] = b.e[j - 1]; --j; }
        b.e[j] = v;
    }
    int64_t n = 0, nacc = 0;
    for (int64_t i = 0; i < b.n; ++i) {
        int64_t ntot = n + b.e[i].tip->n;
        if (ntot > k && n > 0) break;
        n = ntot; nacc = i + 1;
    }
    /* acc = xsh_last <> ... <> xsh_first */
    int64_t m = 0;
    for (int64_t i = nacc - 1; i >= 0; --i)
        for (int64_t e = 0; e < b.e[i].tip->n; ++e) {
            if (m < cap) { ids[m] = b.e[i].tip->ids[e]; dist[m] = point_dist(f, b.e[i].tip->ids[e], q); }
            ++m;
        }
    free(b.e);
    return m;
}
int64_t orc_knn_h(const orc_forest* f, const double* q, int32_t k, double* dist, uint32_t* ids, int64_t cap) {
    qref r = {q, 0, NULL, NULL};
    return knn_h_q(f, &r, k, dist, ids, cap);
}
int64_t orc_knn_h_sq(const orc_forest* f, int64_t qnz, const int32_t* qidx, const double* qval, int32_t k, double* dist, uint32_t* ids, int64_t cap) {
    qref r = {NULL, qnz, qidx, qval};
    return knn_h_q(f, &r, k, dist, ids, cap);
}

static int cmp_u32(const void* a, const void* b) { uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b; return x < y ? -1 : x > y; }

/* recallWith / recallWith1, RPTree.hs:265-282.  `points tt` = leaves left to right; Data.List.sortBy is
 * a stable merge sort.  Sets are over row ids here (the reference's Set is over Embed values, which
 * collapses exact duplicate vectors carrying equal payloads). */
static double recall_q(const orc_forest* f, const qref* q, int32_t k) {
    double sum = 0.0;
    pk* a = (pk*)malloc(sizeof(pk) * (size_t)(f->n > 0 ? f->n : 1));
    pk* tmp = (pk*)malloc(sizeof(pk) * (size_t)(f->n > 0 ? f->n : 1));
    int64_t nn_cap = 0; int64_t* child = NULL; int32_t* depth = NULL; double *thr = NULL, *mlo = NULL, *mhi = NULL; int64_t *ss = NULL, *sz = NULL;
    uint32_t* perm = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(f->n > 0 ? f->n : 1));
    for (int32_t t = 0; t < f->T; ++t) {
        int64_t nn = count_nodes(f->roots[t]);
        if (nn > nn_cap) {
            nn_cap = nn;
            child = (int64_t*)realloc(child, sizeof(int64_t) * (size_t)nn); depth = (int32_t*)realloc(depth, sizeof(int32_t) * (size_t)nn);
            thr = (double*)realloc(thr, sizeof(double) * (size_t)nn); mlo = (double*)realloc(mlo, sizeof(double) * (size_t)nn); mhi = (double*)realloc(mhi, sizeof(double) * (size_t)nn);
            ss = (int64_t*)realloc(ss, sizeof(int64_t) * (size_t)nn); sz = (int64_t*)realloc(sz, sizeof(int64_t) * (size_t)nn);
        }
        orc_tree_export(f, t, child, depth, thr, mlo, mhi, ss, sz, perm);
        int64_t np = count_points(f->roots[t]);
        for (int64_t i = 0; i < np; ++i) { a[i].id = perm[i]; a[i].key = point_dist(f, perm[i], q); }
        msort(a, tmp, np);
        int64_t kk = np < k ? np : k;
        uint32_t* top = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(kk > 0 ? kk : 1));
        for (int64_t i = 0; i < kk; ++i) top[i] = a[i].id;
        qsort(top, (size_t)kk, sizeof(uint32_t), cmp_u32);
        idbuf b = {0};
        cand_go(f, t, 0, f->roots[t], q, &b);
        if (b.n > 0) qsort(b.ids, (size_t)b.n, sizeof(uint32_t), cmp_u32);
        int64_t hit = 0, i = 0, j = 0;
        uint32_t last = 0; int have_last = 0;
        while (i < b.n && j < kk) {
            if (b.ids[i] < top[j]) ++i;
            else if (b.ids[i] > top[j]) ++j;
            else { if (!have_last || last != top[j]) { ++hit; last = top[j]; have_last = 1; } ++i; ++j; }
        }
        sum = sum + (double)hit / (double)k;
        free(top); free(b.ids);
    }
    free(a); free(tmp); free(child); free(depth); free(thr); free(mlo); free(mhi); free(ss); free(sz); free(perm);
    return sum / (double)f->T;
}

/* recallWith with the brute-force distances evaluated ONCE instead of once per tree (same value as recall_q).  In a batch
 * forest every tree holds all n rows, so the per-tree truth sets of RPTree.hs:279 differ only in how a tie AT the k-th
 * distance is cut (each tree's own leaf order); when such a tie exists, or a tree does not hold all rows (streaming build
 * with dropped subtrees), the query is answered by recall_q itself. */
static double recall_shared_q(const orc_forest* f, const qref* q, int32_t k) {
    int64_t n = f->n;
    if (n < 1 || k < 1) return recall_q(f, q, k);
    for (int32_t t = 0; t < f->T; ++t) if (count_points(f->roots[t]) != n) return recall_q(f, q, k);
    double* dd = (double*)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) dd[i] = point_dist(f, (uint32_t)i, q);
    int64_t kk = n < k ? n : k;
    /* k-th smallest distance: keep the kk smallest seen so far in a sorted buffer */
    double* best = (double*)malloc(sizeof(double) * (size_t)kk);
    int64_t nb = 0;
    for (int64_t i = 0; i < n; ++i) {
        double v = dd[i];
        if (nb == kk && !(v < best[kk - 1])) continue;
        int64_t p = nb < kk ? nb++ : kk - 1;
        while (p > 0 && v < best[p - 1]) { best[p] = best[p - 1]; --p; }
        best[p] = v;
    }
    double tau = best[kk - 1];
    free(best);
    int64_t nle = 0;
    for (int64_t i = 0; i < n; ++i) if (dd[i] <= tau) ++nle;
    if (nle != kk) { free(dd); return recall_q(f, q, k); }          /* tie across the cut: tree order decides */
    uint32_t* top = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)kk);
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) if (dd[i] <= tau) top[m++] = (uint32_t)i;   /* ascending row id */
    free(dd);
    double sum = 0.0;
    for (int32_t t = 0; t < f->T; ++t) {
        idbuf b = {0};
        cand_go(f, t, 0, f->roots[t], q, &b);
        if (b.n > 0) qsort(b.ids, (size_t)b.n, sizeof(uint32_t), cmp_u32);
        int64_t hit = 0, i = 0, j = 0;
        uint32_t last = 0; int have_last = 0;
        while (i < b.n && j < kk) {
            if (b.ids[i] < top[j]) ++i;
            else if (b.ids[i] > top[j]) ++j;
            else { if (!have_last || last != top[j]) { ++hit; last = top[j]; have_last = 1; } ++i; ++j; }
        }
        sum = sum + (double)hit / (double)k;
        free(b.ids);
    }
    free(top);
    return sum / (double)f->T;
}
double orc_recall_shared(const orc_forest* f, const double* q, int32_t k) {
    qref r = {q, 0, NULL, NULL};
    return recall_shared_q(f, &r, k);
}
double orc_recall(const orc_forest* f, const double* q, int32_t k) {
    qref r = {q, 0, NULL, NULL};
    return recall_q(f, &r, k);
}
double orc_recall_sq(const orc_forest* f, int64_t qnz, const int32_t* qidx, const double* qval, int32_t k) {
    qref r = {NULL, qnz, qidx, qval};
    return recall_q(f, &r, k);
}

void orc_brute_knn(const double* X, int64_t n, int32_t d, const double* q, int32_t k, double* dist, uint32_t* ids) {
    pk* a = (pk*)malloc(sizeof(pk) * (size_t)(n > 0 ? n : 1));
    pk* tmp = (pk*)malloc(sizeof(pk) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) { a[i].id = (uint32_t)i; a[i].key = orc_metric_dd_l2(X + i * d, q, d); }
    msort(a, tmp, n);
    for (int64_t i = 0; i < k && i < n; ++i) { dist[i] = a[i].key; ids[i] = a[i].id; }
    free(a); free(tmp);
}
