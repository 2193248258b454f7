// extern "C" surface of libb200sd.so (declared in include/b200sd.h).
#include "../../include/b200sd.h"

#include "igemm.cuh"

using namespace b2;

extern "C" {

const char* b2sd_last_error(void) { return b2_last_error(); }
int b2sd_version(void) { return 1; }

static ActView to_view(const b2sd_act_view& v) {
    ActView a;
    a.ptr = reinterpret_cast<const __half*>(v.ptr);
    a.N = v.n; a.H = v.h; a.W = v.w; a.C = v.c; a.ld = v.ld;
    return a;
}

int b2sd_op_igemm(const b2sd_igemm_desc* d, void* stream) {
    if (!d) {
        b2_set_error("b2sd_op_igemm: null desc");
        return -1;
    }
    IgemmDesc g{};
    g.nseg = d->nseg;
    for (int s = 0; s < d->nseg && s < IG_MAX_SRC; ++s) {
        g.src[s] = to_view(d->src[s]);
        g.ntap[s] = d->ntap[s];
    }
    g.w = reinterpret_cast<const __half*>(d->w);
    g.w_rows = d->w_rows;
    g.w_ld = d->w_ld;
    g.stride = d->stride;
    g.Nb = d->nb; g.Ho = d->ho; g.Wo = d->wo;
    g.BN = d->bn;
    g.splits = d->splits;
    g.partial = reinterpret_cast<float*>(d->partial);
    g.epi.out = reinterpret_cast<__half*>(d->out);
    g.epi.ldc = d->ldc;
    g.epi.colbias = d->colbias;
    g.epi.colbias_bstride = d->colbias_bstride;
    g.epi.res = reinterpret_cast<const __half*>(d->res);
    g.epi.ldr = d->ldr;
    g.epi.acc_scale = d->acc_scale;
    g.epi.res_scale = d->res_scale;
    g.epi.flags = d->flags & (IG_RELU | IG_GEGLU);
    g.epi.n_valid = d->n_valid;
    IgemmPlan plan;
    if (igemm_plan(g, &plan)) return -1;
    return igemm_launch(plan, reinterpret_cast<cudaStream_t>(stream));
}

uint64_t b2sd_igemm_partial_floats(int splits, int64_t rows_total, int n_valid) {
    return igemm_partial_floats(splits, rows_total, n_valid);
}

}  // extern "C"
