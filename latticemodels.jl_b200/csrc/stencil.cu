// Registry of the compiled stencil patterns and register-tile variants (kernels: stencil.cuh,
// instantiated per pattern in stencil_k<id>.cu).
#include "stencil.cuh"
#include <cstdlib>

namespace lm {

// compiled sparsity patterns (|d| <= 1 cell offsets); a detected pattern runs on the smallest
// compiled superset (missing entries carry the value 0)
//   id 0  square / rectangular NN + on-site (RC = 1)
//   id 1  any RC = 1 pattern with |d| <= 1 (NNN square, triangular)
//   id 2  honeycomb NN + on-site (RC = 2)
//   id 3  QWZ-type: 2 orbitals, full 2x2 blocks on-site and to the 4 nearest cells
//   id 4  Haldane: honeycomb NN + NNN + on-site
//   id 5  any RC = 2 pattern with |d| <= 1 (honeycomb with third neighbours, 2-orbital models with diagonal hops)
//   id 6  kagome NN + on-site (RC = 3)
//   id 7  kagome NN + second neighbours + on-site (RC = 3)
//   id 8  Kane-Mele: spin-1/2 honeycomb, spin-diagonal NN + second neighbours + on-site diagonal (RC = 4)
//   id 9  QWZ with a diagonal on-site term (what `qwz` builds): 9 entries per row instead of the 10 of id 3

#define LM_ST_DESC(id, name) {StPat<id>::rc, StPat<id>::mask, StPat<id>::imag, st_width<StPat<id>::rc>(StPat<id>::mask), name}
static const StencilDesc g_desc[] = {
    LM_ST_DESC(0, "square-nn"), LM_ST_DESC(1, "rc1-full"), LM_ST_DESC(2, "honeycomb-nn"),
    LM_ST_DESC(3, "qwz"), LM_ST_DESC(4, "haldane"), LM_ST_DESC(5, "rc2-full"),
    LM_ST_DESC(6, "kagome-nn"), LM_ST_DESC(7, "kagome-nnn"), LM_ST_DESC(8, "kanemele"), LM_ST_DESC(9, "qwz-diag"),
};
static_assert(sizeof(g_desc) / sizeof(g_desc[0]) == LM_ST_NPAT, "pattern table out of step with StPat<>");
int stencil_count() { return LM_ST_NPAT; }
const StencilDesc& stencil_desc(int id) { return id >= LM_ST_RTC_BASE ? *stencil_rtc_desc(id) : g_desc[id]; }
int stencil_find(int rc, const st_mask_t& mask) {
    static const long skip = getenv("LM_STENCIL_SKIP") ? strtol(getenv("LM_STENCIL_SKIP"), nullptr, 0) : 0;     // debug / A-B runs: bit i hides pattern i
    int best = -1;
    for (int i = 0; i < stencil_count(); ++i) {
        if (g_desc[i].rc != rc || !st_covers(g_desc[i].mask, mask) || ((skip >> i) & 1)) continue;
        if (best < 0 || g_desc[i].sw < g_desc[best].sw) best = i;
    }
    return best;
}

int stencil_diag_slot(int id, int a) {
    const StencilDesc& d = stencil_desc(id);
    const int rc = d.rc;
    int s = 0;
    for (int o = 0; o < 9; ++o) for (int b = 0; b < rc; ++b) {
        const bool set = st_get(d.mask, o * rc * rc + a * rc + b);
        if (o == 4 && b == a) return set ? s : -1;
        if (set) ++s;
    }
    return -1;
}
int stencil_stride(int id, bool c64) { const int sw = stencil_desc(id).sw; return c64 ? ((sw + 1) & ~1) : sw; }
int stencil_rstride(int id, bool c64) { const int sw = stencil_desc(id).sw; return c64 ? ((sw + 3) & ~3) : ((sw + 1) & ~1); }

// register-tile variants: T1 x T2 cells per thread, W1 x W2 warps per CTA, CPT lane elements per
// thread; staged = haloed patch brought into shared memory by TMA bulk copies
struct Variant { int t1, t2, w1, w2, cpt, staged; };
static const Variant g_var[] = {
    {4, 2, 2, 4, 1, 0},   // 0: direct, 8 x 8 cell patch
    {2, 2, 2, 4, 2, 0},   // 1: direct, 4 x 8, two lane elements per thread
    {4, 2, 2, 2, 1, 1},   // 2: staged, 8 x 4 patch, 128 threads            (RC = 2 default)
    {4, 2, 2, 4, 1, 1},   // 3: staged, 8 x 8 patch, 256 threads
    {2, 2, 2, 2, 2, 1},   // 4: staged, 4 x 4 patch, two lane elements
    {2, 4, 2, 2, 1, 1},   // 5: staged, 4 x 8 patch
    {2, 2, 4, 2, 1, 1},   // 6: staged, 8 x 4 patch of 2 x 2 tiles, 256 threads
    {4, 4, 2, 2, 1, 1},   // 7: staged, 8 x 8 patch, 128 threads              (RC = 1 default)
    {4, 4, 2, 2, 1, 0},   // 8: direct, 8 x 8 patch                           (RC = 1)
    {4, 2, 2, 4, 2, 1},   // 9: staged, 8 x 8 patch, two lane elements        (RC = 1)
    {2, 2, 4, 2, 1, 2},   // 10: streaming (3-stage TMA pipeline over a column group), 8 x 4 patch, 256 threads
    {4, 2, 2, 2, 1, 2},   // 11: streaming, 8 x 4 patch, 128 threads
    {4, 4, 2, 2, 1, 2},   // 12: streaming, 8 x 8 patch, 128 threads          (RC = 1 only)
    {4, 2, 1, 2, 1, 1},   // 13: staged, 4 x 4 patch, 64 threads  (RC = 2)    - more, smaller CTAs per SM: finer overlap of fill and compute
    {4, 2, 2, 1, 1, 1},   // 14: staged, 8 x 2 patch, 64 threads  (RC = 2)
    {4, 2, 1, 3, 1, 1},   // 15: staged, 4 x 6 patch, 96 threads  (RC = 2)
    {4, 4, 1, 2, 1, 1},   // 16: staged, 4 x 8 patch, 64 threads  (RC = 1)
    {4, 4, 2, 1, 1, 1},   // 17: staged, 8 x 4 patch, 64 threads  (RC = 1)
    {0, 0, 1, 1, 1, 1},   // 18: staged, one warp per CTA: the patch is the tile (4 x 2 cells RC = 2, 4 x 4 RC = 1)
    {2, 2, 2, 2, 1, 1},   // 19: staged, 4 x 4 patch of 2 x 2 tiles, 128 threads                (RC = 3 / 4 default: 12 / 16 rows per thread)
};
int stencil_num_variants() { return (int)(sizeof(g_var) / sizeof(g_var[0])); }
void stencil_variant_shape(int v, int rc, int* P1, int* P2, int* cpt, int* staged) {
    int t1 = g_var[v].t1, t2 = g_var[v].t2;
    if (t1 == 0) { t1 = 4; t2 = rc == 1 ? 4 : 2; }             // variant 18: the default tile of the pattern
    *P1 = g_var[v].w1 * t1; *P2 = g_var[v].w2 * t2; *cpt = g_var[v].cpt; *staged = g_var[v].staged;
}

// CTAs of k_apply_stencil_tma resident per SM (host restatement of st_tma_smem / st_tma_blocks)
int stencil_resident_ctas(int id, int v, bool c64) {
    Variant q = g_var[v];
    const int rc = stencil_desc(id).rc, sw = stencil_stride(id, c64);
    if (q.t1 == 0) { q.t1 = 4; q.t2 = rc == 1 ? 4 : 2; }
    const int P1 = q.w1 * q.t1, P2 = q.w2 * q.t2;
    const size_t smem = (size_t)(P1 + 2) * (P2 + 2) * rc * 32 * q.cpt * 16 + (size_t)P1 * P2 * rc * sw * (c64 ? 8 : 16);
    const int by_smem = (int)((227 * 1024) / (smem + 1024 + 64));
    const int by_regs = st_min_blocks(q.t1 * q.t2 * rc * q.cpt * 4, 32 * q.w1 * q.w2);
    const int m = by_smem < by_regs ? by_smem : by_regs;
    return m < 1 ? 1 : m;
}

#define LM_ST_EACH(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9)
#define LM_ST_DECL(i) int stencil_launch_##i(int, bool, int, const StencilArgs&, const CUtensorMap&, dim3, cudaStream_t); \
    int stencil_observe_##i(bool, const StencilObsArgs&, const CUtensorMap&, unsigned, cudaStream_t); void stencil_obs_shape_##i(int*, int*, int*);
LM_ST_EACH(LM_ST_DECL)
#undef LM_ST_DECL

int stencil_launch(int id, int variant, bool c64, int mode, const StencilArgs& a, const CUtensorMap& tmx, dim3 grid, cudaStream_t s) {
    if (id >= LM_ST_RTC_BASE) return stencil_rtc_launch(id, c64, mode, a, tmx, grid, s);       // one shape per rows-per-cell: `variant` is ignored
    switch (id) {
#define LM_ST_CASE(i) case i: return stencil_launch_##i(variant, c64, mode, a, tmx, grid, s);
    LM_ST_EACH(LM_ST_CASE)
#undef LM_ST_CASE
    default: return -1;
    }
}
int stencil_observe(int id, bool c64, const StencilObsArgs& a, const CUtensorMap& tmx, unsigned grid, cudaStream_t s) {
    if (id >= LM_ST_RTC_BASE) return stencil_rtc_observe(id, c64, a, tmx, grid, s);
    switch (id) {
#define LM_ST_CASE(i) case i: return stencil_observe_##i(c64, a, tmx, grid, s);
    LM_ST_EACH(LM_ST_CASE)
#undef LM_ST_CASE
    default: return -1;
    }
}
void stencil_obs_shape(int id, int* P1, int* P2, int* nf) {
    if (id >= LM_ST_RTC_BASE) {
        const StencilDesc& d = stencil_desc(id);
        int t1, t2, w1, w2;
        *nf = stencil_rtc_nfwd(d.rc, d.mask);
        if (!stencil_rtc_obs_shape(d.rc, *nf, &t1, &t2, &w1, &w2)) { *P1 = *P2 = 0; return; }       // too many forward entries per cell: ELL-plan observables
        *P1 = w1 * t1; *P2 = w2 * t2;
        return;
    }
    switch (id) {
#define LM_ST_CASE(i) case i: stencil_obs_shape_##i(P1, P2, nf); break;
    LM_ST_EACH(LM_ST_CASE)
#undef LM_ST_CASE
    default: *P1 = *P2 = *nf = 0; break;
    }
}

}  // namespace lm
