// C-ABI entry points that expose single kernels for the per-kernel parity tests (tests/test_kernels_gpu.py).
#include "conv_plan.cuh"

using namespace dsb;

static int num_sms_cached() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

static int g_test_two_cta = 0;
// -1: never use CTA pairs, 0: automatic, 1: always (for the per-kernel tests)
extern "C" void dsb_test_set_two_cta(int mode) { g_test_two_cta = mode; }

extern "C" int dsb_test_conv(int kind, int F, int H, int W, int Cin, int N, int dilation, int T, int kt,
                             const void* A, const void* Wt, const float* scale, const float* shift,
                             const float* rowbias, const float* residual, int act, float* out_f32, void* out_bf16,
                             int out_fmul, int out_fadd, const float* head_w, float head_b, float* out_head,
                             void* stream) {
    ConvOp op;
    memset(&op, 0, sizeof(op));
    op.kind = kind; op.F = F; op.H = H; op.W = W; op.Cin = Cin; op.N = N; op.dilation = dilation; op.T = T; op.kt = kt;
    op.A = (const bf16*)A; op.Wt = (const bf16*)Wt;
    op.scale = scale; op.shift = shift; op.rowbias = rowbias; op.residual = residual; op.act = act;
    op.out_f32 = out_f32; op.out_bf16 = (bf16*)out_bf16; op.out_fmul = out_fmul; op.out_fadd = out_fadd;
    op.head_w = head_w; op.head_b = head_b; op.out_head = out_head;
    op.two_cta = g_test_two_cta;
    ConvLaunch l;
    int r = conv_lower(op, &l);
    if (r) return r;
    return conv_run(l, num_sms_cached(), (cudaStream_t)stream);
}
