/*
 * ovdet_oracle.c -- CPU restatement of the reference's detection-geometry path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / CPU baseline.  The
 * product path (open-vocabulary-3d-object-detection_b200/) never links or
 * imports it and fails loudly when its CUDA library is missing.
 *
 * Parity status: PINNED against outputs of the reference itself, run in the
 * build container (tests/golden/make_golden.py imports /root/reference and the
 * Cython extension compiled by oracle/build_ref.py; fixtures are committed under
 * tests/golden/).  The reference has no tests of its own.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the reference checkout).  Compile with -ffp-contract=off: the reference
 * evaluates each arithmetic op separately (Python floats / eager torch / numpy),
 * so no fused multiply-add may be introduced here.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXV 32

/* ------------------------------------------------------------------------- */
/* Sutherland-Hodgman clip of `subj` (ns vertices) by convex `clip` (nc),     */
/* vertices interleaved x,y.  Restates utils/box_intersection.pyx:27-70       */
/* (= utils/box_util.py:34-81 and :404-440): strict `>` inside test           */
/* (pyx:23-24), intersection formula pyx:13-19, early exit on empty output.   */
/* Instantiated for double (Python-float arithmetic of the Cython/numpy       */
/* paths) and float (fp32 torch path).                                        */
/* ------------------------------------------------------------------------- */
#define DEFINE_SH_CLIP(NAME, T)                                                \
static int NAME(const T *subj, int ns, const T *clip, int nc, T *out)          \
{                                                                              \
    T a[2 * MAXV], b[2 * MAXV];                                                \
    T *cur = a, *nxt = b;                                                      \
    int n = ns;                                                                \
    memcpy(cur, subj, sizeof(T) * 2 * (size_t)ns);                             \
    T c1x = clip[2 * (nc - 1)], c1y = clip[2 * (nc - 1) + 1];                  \
    for (int ci = 0; ci < nc; ++ci) {                                          \
        const T c2x = clip[2 * ci], c2y = clip[2 * ci + 1];                    \
        const T ex_ = c2x - c1x, ey_ = c2y - c1y;                              \
        int m = 0;                                                             \
        T sx = cur[2 * (n - 1)], sy = cur[2 * (n - 1) + 1];                    \
        for (int i = 0; i < n; ++i) {                                          \
            const T px = cur[2 * i], py = cur[2 * i + 1];                      \
            const int e_in = ex_ * (py - c1y) > ey_ * (px - c1x);              \
            const int s_in = ex_ * (sy - c1y) > ey_ * (sx - c1x);              \
            if (e_in != s_in && m < MAXV) {                                    \
                const T dcx = c1x - c2x, dcy = c1y - c2y;                      \
                const T dpx = sx - px, dpy = sy - py;                          \
                const T n1 = c1x * c2y - c1y * c2x;                            \
                const T n2 = sx * py - sy * px;                                \
                const T n3 = (T)1.0 / (dcx * dpy - dcy * dpx);                 \
                nxt[2 * m] = (n1 * dpx - n2 * dcx) * n3;                       \
                nxt[2 * m + 1] = (n1 * dpy - n2 * dcy) * n3;                   \
                ++m;                                                           \
            }                                                                  \
            if (e_in && m < MAXV) {                                            \
                nxt[2 * m] = px;                                               \
                nxt[2 * m + 1] = py;                                           \
                ++m;                                                           \
            }                                                                  \
            sx = px;                                                           \
            sy = py;                                                           \
        }                                                                      \
        c1x = c2x;                                                             \
        c1y = c2y;                                                             \
        T *t = cur; cur = nxt; nxt = t;                                        \
        n = m;                                                                 \
        if (n == 0) break;                                                     \
    }                                                                          \
    memcpy(out, cur, sizeof(T) * 2 * (size_t)n);                               \
    return n;                                                                  \
}

DEFINE_SH_CLIP(sh_clip_f64, double)
DEFINE_SH_CLIP(sh_clip_f32, float)

/* Public wrappers for tests: clip and return vertex count. */
int oracle_polygon_clip_f64(const double *subj, int ns, const double *clip, int nc, double *out)
{
    return sh_clip_f64(subj, ns, clip, nc, out);
}
int oracle_polygon_clip_f32(const float *subj, int ns, const float *clip, int nc, float *out)
{
    return sh_clip_f32(subj, ns, clip, nc, out);
}

/* Shoelace area 0.5*|x . roll(y,1) - y . roll(x,1)| in double.
 * utils/box_util.py:84-86 (poly_area).  Also stands in for
 * scipy.spatial.ConvexHull(...).volume at box_util.py:96 -- the clip of two
 * convex polygons is convex, so hull area == polygon area (Qhull, scipy 1.18.1
 * as installed, is the third-party dependency; it is not restated further). */
double oracle_poly_area_f64(const double *p, int n)
{
    double d1 = 0.0, d2 = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + n - 1) % n; /* roll(.,1)[i] = [i-1] */
        d1 += p[2 * i] * p[2 * j + 1];
        d2 += p[2 * i + 1] * p[2 * j];
    }
    return 0.5 * fabs(d1 - d2);
}

/* ------------------------------------------------------------------------- */
/* utils/box_intersection.pyx:166-198 -- the Cython hot loop, as shipped.     */
/* k2_loop is `rect2.shape[2]` (pyx:180), i.e. 4 for the shipped code: only   */
/* columns k2 < min(k2_loop, nums_k2[b]) are ever clipped.  Clip arithmetic   */
/* is Python-float (fp64) on fp32 inputs; the polygon is cast to fp32         */
/* (pyx:196-197) and the shoelace uses np.dot on fp32 vectors, which          */
/* (numpy 2.3.5 / OpenBLAS, probed) rounds each product to fp32, accumulates  */
/* in double and rounds the sum to fp32; 0.5*abs(.) stays fp32 (pyx:198).     */
/* inter_areas is written in place, untouched where nothing is clipped.       */
/* ------------------------------------------------------------------------- */
static float shoelace_npdot_f32(const float *xs, const float *ys, int n)
{
    double d1 = 0.0, d2 = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + n - 1) % n;
        d1 += (double)(float)(xs[i] * ys[j]);
        d2 += (double)(float)(ys[i] * xs[j]);
    }
    const float f1 = (float)d1, f2 = (float)d2;
    return 0.5f * fabsf(f1 - f2);
}

void oracle_box_intersection(const float *rect1, const float *rect2,
                             const float *non_rot_inter_areas, const int32_t *nums_k2,
                             float *inter_areas, int approximate,
                             int B, int K1, int K2, int k2_loop)
{
    for (int b = 0; b < B; ++b)
        for (int k1 = 0; k1 < K1; ++k1)
            for (int k2 = 0; k2 < k2_loop && k2 < K2; ++k2) {
                if (k2 >= nums_k2[b]) break;
                const size_t o = ((size_t)b * K1 + k1) * K2 + k2;
                if (approximate && non_rot_inter_areas[o] == 0.0f) continue;
                const float *r1 = rect1 + ((size_t)b * K1 + k1) * 8;
                const float *r2 = rect2 + ((size_t)b * K2 + k2) * 8;
                double s[8], c[8], poly[2 * MAXV];
                for (int i = 0; i < 8; ++i) { s[i] = r1[i]; c[i] = r2[i]; }
                const int n = sh_clip_f64(s, 4, c, 4, poly);
                if (n > 0) {
                    float xs[MAXV], ys[MAXV];
                    for (int i = 0; i < n; ++i) { xs[i] = (float)poly[2 * i]; ys[i] = (float)poly[2 * i + 1]; }
                    inter_areas[o] = shoelace_npdot_f32(xs, ys, n);
                }
            }
}

/* ------------------------------------------------------------------------- */
/* generalized_box3d_iou: utils/box_util.py:517-618 (torch path, clip_mode 0) */
/* and :624-714 (Cython-backed path, clip_mode 1).  corners [B,K,8,3] fp32.   */
/*   height      box_util.py:544-547 / :651-653                               */
/*   BEV rects   corners [3,2,1,0], columns (0,2)   :550-555 / :656-661        */
/*   prefilter   rect points 1 and 3 as lt/rb       :557-564 / :663-670        */
/*   enclosing   axis-aligned box of all 16 corners :466-514                   */
/*   volumes     edge lengths sqrt(clamp(.,1e-6))   :443-463, clamp 1e-8 :568  */
/*   good_boxes  :574 ; union/iou/giou/masks        :600-618 / :700-714        */
/* Switches mirror the reference variants (SURVEY.md "three quirks"):         */
/*   prefilter  skip pairs whose axis-aligned BEV overlap is 0 (:587-588)     */
/*   k2_cap     >0: clip only columns < k2_cap (the pyx:180 bug, =4)           */
/*   enclosing_hull: not here (see oracle/__init__.py, scipy ConvexHull).     */
/* ------------------------------------------------------------------------- */
static float box_vol_f32(const float *c)
{
    const float EPS = 1e-6f;
    float e[3];
    const int pa[3] = {0, 1, 0}, pb[3] = {1, 2, 4};
    for (int t = 0; t < 3; ++t) {
        const float dx = c[3 * pa[t]] - c[3 * pb[t]];
        const float dy = c[3 * pa[t] + 1] - c[3 * pb[t] + 1];
        const float dz = c[3 * pa[t] + 2] - c[3 * pb[t] + 2];
        float s = (dx * dx + dy * dy) + dz * dz;
        if (s < EPS) s = EPS;
        e[t] = sqrtf(s);
    }
    return (e[0] * e[1]) * e[2];
}

void oracle_giou3d(const float *corners1, const float *corners2, const int64_t *nums_k2,
                   int B, int K1, int K2, int rotated, int inter_only, int clip_mode,
                   int prefilter, int k2_cap, float *out)
{
    const float EPS = 1e-8f;
    for (int b = 0; b < B; ++b) {
        const int nk2 = nums_k2 ? (int)nums_k2[b] : K2;
        for (int k1 = 0; k1 < K1; ++k1) {
            const float *c1 = corners1 + ((size_t)b * K1 + k1) * 24;
            float r1[8];
            for (int i = 0; i < 4; ++i) { r1[2 * i] = c1[3 * (3 - i)]; r1[2 * i + 1] = c1[3 * (3 - i) + 2]; }
            float mn1[3], mx1[3];
            for (int a = 0; a < 3; ++a) {
                mn1[a] = mx1[a] = c1[a];
                for (int i = 1; i < 8; ++i) { mn1[a] = fminf(mn1[a], c1[3 * i + a]); mx1[a] = fmaxf(mx1[a], c1[3 * i + a]); }
            }
            float v1 = box_vol_f32(c1);
            if (v1 < EPS) v1 = EPS;
            for (int k2 = 0; k2 < K2; ++k2) {
                const float *c2 = corners2 + ((size_t)b * K2 + k2) * 24;
                float r2[8];
                for (int i = 0; i < 4; ++i) { r2[2 * i] = c2[3 * (3 - i)]; r2[2 * i + 1] = c2[3 * (3 - i) + 2]; }
                const int valid = k2 < nk2;
                float h = fminf(c1[1], c2[1]) - fmaxf(c1[13], c2[13]);
                if (h < 0.0f) h = 0.0f;
                float w0 = fminf(r1[6], r2[6]) - fmaxf(r1[2], r2[2]);
                float w1 = fminf(r1[7], r2[7]) - fmaxf(r1[3], r2[3]);
                if (w0 < 0.0f) w0 = 0.0f;
                if (w1 < 0.0f) w1 = 0.0f;
                float non_rot = w0 * w1;
                if (!valid) non_rot = 0.0f;
                float d[3];
                for (int a = 0; a < 3; ++a) {
                    float mn = mn1[a], mx = mx1[a];
                    for (int i = 0; i < 8; ++i) { mn = fminf(mn, c2[3 * i + a]); mx = fmaxf(mx, c2[3 * i + a]); }
                    d[a] = fabsf(mx - mn);
                }
                const float encl = (d[0] * d[1]) * d[2];
                float v2 = box_vol_f32(c2);
                if (v2 < EPS) v2 = EPS;
                const float sum_vols = v1 + v2;
                const float good = (encl > 2 * EPS) && (sum_vols > 4 * EPS) ? 1.0f : 0.0f;
                float area = 0.0f;
                if (rotated) {
                    int do_clip = valid && !(prefilter && non_rot == 0.0f);
                    if (k2_cap > 0 && k2 >= k2_cap) do_clip = 0;
                    if (do_clip) {
                        if (clip_mode == 1) {
                            double s[8], c[8], poly[2 * MAXV];
                            for (int i = 0; i < 8; ++i) { s[i] = r1[i]; c[i] = r2[i]; }
                            const int n = sh_clip_f64(s, 4, c, 4, poly);
                            if (n > 0) {
                                float xs[MAXV], ys[MAXV];
                                for (int i = 0; i < n; ++i) { xs[i] = (float)poly[2 * i]; ys[i] = (float)poly[2 * i + 1]; }
                                area = shoelace_npdot_f32(xs, ys, n);
                            }
                        } else {
                            float poly[2 * MAXV];
                            const int n = sh_clip_f32(r1, 4, r2, 4, poly);
                            if (n > 0) { /* box_util.py:591-598: abs(dot - dot) then *0.5 */
                                float d1 = 0.0f, d2 = 0.0f;
                                for (int i = 0; i < n; ++i) {
                                    const int j = (i + n - 1) % n;
                                    d1 += poly[2 * i] * poly[2 * j + 1];
                                    d2 += poly[2 * i + 1] * poly[2 * j];
                                }
                                area = fabsf(d1 - d2) * 0.5f;
                            }
                        }
                    }
                } else {
                    area = non_rot;
                }
                const float inter = area * h;
                float *o = out + ((size_t)b * K1 + k1) * K2 + k2;
                if (inter_only) { *o = inter; continue; }
                float uni = sum_vols - inter;
                if (uni < EPS) uni = EPS;
                const float iou = inter / uni;
                const float second = -(1.0f - uni / encl);
                float g = iou + second;
                g *= good;
                if (nums_k2) g *= valid ? 1.0f : 0.0f;
                *o = g;
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* box3d_iou: utils/box_util.py:116-141 (fp64, no prefilter, no clamps).      */
/* Returns iou; *iou2d gets the BEV IoU.  A degenerate clip result (<3 pts or */
/* zero area), where the reference's Qhull call would raise, gives area 0.    */
/* ------------------------------------------------------------------------- */
static double box_vol_f64(const double *c)
{
    double e[3];
    const int pa[3] = {0, 1, 0}, pb[3] = {1, 2, 4};
    for (int t = 0; t < 3; ++t) {
        const double dx = c[3 * pa[t]] - c[3 * pb[t]];
        const double dy = c[3 * pa[t] + 1] - c[3 * pb[t] + 1];
        const double dz = c[3 * pa[t] + 2] - c[3 * pb[t] + 2];
        e[t] = sqrt((dx * dx + dy * dy) + dz * dz);
    }
    return (e[0] * e[1]) * e[2];
}

double oracle_box3d_iou(const double *c1, const double *c2, double *iou2d)
{
    double r1[8], r2[8], poly[2 * MAXV];
    for (int i = 0; i < 4; ++i) {
        r1[2 * i] = c1[3 * (3 - i)]; r1[2 * i + 1] = c1[3 * (3 - i) + 2];
        r2[2 * i] = c2[3 * (3 - i)]; r2[2 * i + 1] = c2[3 * (3 - i) + 2];
    }
    const double a1 = oracle_poly_area_f64(r1, 4), a2 = oracle_poly_area_f64(r2, 4);
    const int n = sh_clip_f64(r1, 4, r2, 4, poly);
    const double ia = n >= 3 ? oracle_poly_area_f64(poly, n) : 0.0;
    if (iou2d) *iou2d = ia / (a1 + a2 - ia);
    const double ymax = fmin(c1[1], c2[1]);
    const double ymin = fmax(c1[13], c2[13]);
    const double iv = ia * fmax(0.0, ymax - ymin);
    const double v1 = box_vol_f64(c1), v2 = box_vol_f64(c2);
    return iv / (v1 + v2 - iv);
}

/* Pairwise exact IoU matrix, dets [nd,8,3] x gts [ng,8,3] given as fp32      */
/* (eval_det.py:120,122 cast fp32 boxes to fp64 before box3d_iou).            */
void oracle_box3d_iou_matrix(const float *dets, int nd, const float *gts, int ng, double *out)
{
    for (int i = 0; i < nd; ++i) {
        double a[24];
        for (int t = 0; t < 24; ++t) a[t] = dets[(size_t)i * 24 + t];
        for (int j = 0; j < ng; ++j) {
            double b[24];
            for (int t = 0; t < 24; ++t) b[t] = gts[(size_t)j * 24 + t];
            out[(size_t)i * ng + j] = oracle_box3d_iou(a, b, NULL);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Greedy NMS: utils/nms.py:43-76 (2d, dims=2), :79-117 (3d), :120-162        */
/* (3d same-class).  boxes fp64 [K, ncols]: dims*2 coords, score, [cls].      */
/* extra_eps is added to every box volume (3DOVDet_tools box_3d_utils.py:74;  */
/* 0 for utils/nms.py).  Order: ascending sort by score, take from the end;   */
/* the sort here is stable on index, numpy's default argsort is not, so       */
/* results are only defined for tie-free scores (documented).                 */
/* Returns number of picks written to `pick` (pick order = reference's).      */
/* ------------------------------------------------------------------------- */
typedef struct { double s; int i; } sidx_t;
static int cmp_sidx(const void *a, const void *b)
{
    const sidx_t *x = (const sidx_t *)a, *y = (const sidx_t *)b;
    if (x->s < y->s) return -1;
    if (x->s > y->s) return 1;
    return (x->i > y->i) - (x->i < y->i);
}

int oracle_nms(const double *boxes, int K, int ncols, int dims, int samecls,
               double thr, int old_type, double extra_eps, int32_t *pick)
{
    if (K <= 0) return 0;
    const int sc = 2 * dims, cc = 2 * dims + 1;
    sidx_t *ord = (sidx_t *)malloc(sizeof(sidx_t) * (size_t)K);
    double *vol = (double *)malloc(sizeof(double) * (size_t)K);
    char *alive = (char *)malloc((size_t)K);
    for (int k = 0; k < K; ++k) {
        const double *bx = boxes + (size_t)k * ncols;
        ord[k].s = bx[sc]; ord[k].i = k; alive[k] = 1;
        double v = bx[dims] - bx[0];
        for (int a = 1; a < dims; ++a) v = v * (bx[dims + a] - bx[a]);
        vol[k] = v + extra_eps;
    }
    qsort(ord, (size_t)K, sizeof(sidx_t), cmp_sidx);
    int np_ = 0;
    for (int p = K - 1; p >= 0; --p) {
        const int i = ord[p].i;
        if (!alive[i]) continue;
        pick[np_++] = i;
        alive[i] = 0;
        const double *bi = boxes + (size_t)i * ncols;
        for (int q = 0; q < p; ++q) {
            const int j = ord[q].i;
            if (!alive[j]) continue;
            const double *bj = boxes + (size_t)j * ncols;
            double inter = 1.0;
            for (int a = 0; a < dims; ++a) {
                double lo = fmax(bi[a], bj[a]), hi = fmin(bi[dims + a], bj[dims + a]);
                double e = fmax(0.0, hi - lo);
                inter = a == 0 ? e : inter * e;
            }
            double o = old_type ? inter / vol[j] : inter / (vol[i] + vol[j] - inter);
            if (samecls) o = o * (bi[cc] == bj[cc] ? 1.0 : 0.0);
            if (o > thr) alive[j] = 0;
        }
    }
    free(ord); free(vol); free(alive);
    return np_;
}

/* ------------------------------------------------------------------------- */
/* VOC AP: utils/eval_det.py:23-54.                                           */
/* ------------------------------------------------------------------------- */
double oracle_voc_ap(const double *rec, const double *prec, int n, int use_07_metric)
{
    if (use_07_metric) {
        double ap = 0.0;
        for (int k = 0; k <= 10; ++k) {
            const double t = k * 0.1; /* np.arange(0.0, 1.1, 0.1)[k] == k*0.1 */
            double p = 0.0; int any = 0;
            for (int i = 0; i < n; ++i)
                if (rec[i] >= t) { if (!any || prec[i] > p) p = prec[i]; any = 1; }
            ap = ap + p / 11.0;
        }
        return ap;
    }
    const int m = n + 2;
    double *mrec = (double *)malloc(sizeof(double) * (size_t)m);
    double *mpre = (double *)malloc(sizeof(double) * (size_t)m);
    mrec[0] = 0.0; mpre[0] = 0.0; mrec[m - 1] = 1.0; mpre[m - 1] = 0.0;
    for (int i = 0; i < n; ++i) { mrec[i + 1] = rec[i]; mpre[i + 1] = prec[i]; }
    for (int i = m - 1; i > 0; --i) if (mpre[i] > mpre[i - 1]) mpre[i - 1] = mpre[i];
    double ap = 0.0;
    for (int i = 0; i < m - 1; ++i)
        if (mrec[i + 1] != mrec[i]) ap += (mrec[i + 1] - mrec[i]) * mpre[i + 1];
    free(mrec); free(mpre);
    return ap;
}

/* ------------------------------------------------------------------------- */
/* eval_det_cls: utils/eval_det.py:66-155 on flattened inputs.                */
/*   det_scene[nd], det_score[nd] (any order), det_box [nd,8,3] fp32          */
/*   gt_scene[ng] (ascending), gt_box [ng,8,3] fp32                           */
/* Sort by -score (stable on index; tie-free scores assumed), per det best    */
/* IoU over the same scene's GT with strict `>` (first max wins, :124-127),   */
/* TP iff ovmax > thr and that GT unclaimed (:130-140).  Writes tp[nd] in     */
/* sorted order, sorted index order[nd]; rec/prec [nd]; returns AP.           */
/* ------------------------------------------------------------------------- */
double oracle_eval_det_cls(const int32_t *det_scene, const float *det_score, const float *det_box, int nd,
                           const int32_t *gt_scene, const float *gt_box, int ng,
                           double ovthresh, int use_07_metric,
                           int32_t *order, double *tp_out, double *rec, double *prec)
{
    sidx_t *ord = (sidx_t *)malloc(sizeof(sidx_t) * (size_t)(nd > 0 ? nd : 1));
    for (int d = 0; d < nd; ++d) { ord[d].s = -(double)det_score[d]; ord[d].i = d; }
    qsort(ord, (size_t)nd, sizeof(sidx_t), cmp_sidx);
    char *claimed = (char *)calloc((size_t)(ng > 0 ? ng : 1), 1);
    double ctp = 0.0, cfp = 0.0;
    const double eps = 2.220446049250313e-16;
    for (int d = 0; d < nd; ++d) {
        const int di = ord[d].i;
        order[d] = di;
        double a[24];
        for (int t = 0; t < 24; ++t) a[t] = det_box[(size_t)di * 24 + t];
        double ovmax = -INFINITY; int jmax = -1;
        for (int j = 0; j < ng; ++j) {
            if (gt_scene[j] != det_scene[di]) continue;
            double b[24];
            for (int t = 0; t < 24; ++t) b[t] = gt_box[(size_t)j * 24 + t];
            const double iou = oracle_box3d_iou(a, b, NULL);
            if (iou > ovmax) { ovmax = iou; jmax = j; }
        }
        double tp = 0.0;
        if (ovmax > ovthresh && !claimed[jmax]) { tp = 1.0; claimed[jmax] = 1; }
        tp_out[d] = tp;
        ctp += tp; cfp += 1.0 - tp;
        rec[d] = ng == 0 ? 0.0 : ctp / (double)ng;
        prec[d] = ctp / fmax(ctp + cfp, eps);
    }
    free(ord); free(claimed);
    return oracle_voc_ap(rec, prec, nd, use_07_metric);
}

/* ------------------------------------------------------------------------- */
/* Axis-aligned 1xN IoU with +eps in the denominator:                         */
/* utils/label_formatter.py:10-64 = 3DOVDet_tools/utils/box_3d_utils.py:3-57. */
/* typ 0 = 'vv' (two corners), 1 = 'cs' (centre + size).  fp64.               */
/* ------------------------------------------------------------------------- */
void oracle_box_3d_iou_aabb(const double *q, const double *k, int n, int kstride, int typ, double eps, double *out)
{
    double ql[3], qh[3];
    for (int a = 0; a < 3; ++a) {
        if (typ == 0) { ql[a] = q[a]; qh[a] = q[3 + a]; }
        else { ql[a] = q[a] - q[3 + a] / 2; qh[a] = q[a] + q[3 + a] / 2; }
    }
    const double qv = ((qh[0] - ql[0]) * (qh[1] - ql[1])) * (qh[2] - ql[2]);
    for (int i = 0; i < n; ++i) {
        const double *b = k + (size_t)i * kstride;
        double kl[3], kh[3];
        for (int a = 0; a < 3; ++a) {
            if (typ == 0) { kl[a] = b[a]; kh[a] = b[3 + a]; }
            else { kl[a] = b[a] - b[3 + a] / 2; kh[a] = b[a] + b[3 + a] / 2; }
        }
        const double kv = ((kh[0] - kl[0]) * (kh[1] - kl[1])) * (kh[2] - kl[2]);
        double inter = 1.0;
        for (int a = 0; a < 3; ++a) {
            const double e = fmax(fmin(qh[a], kh[a]) - fmax(ql[a], kl[a]), 0.0);
            inter = a == 0 ? e : inter * e;
        }
        out[i] = inter / (((qv + kv) - inter) + eps);
    }
}
