// Test-only: compiles sihl_b200/csrc/od_math.h for the HOST so the analytic gradients and the
// CIoU restatement can be checked against torch autograd / torchvision where no GPU exists.
// Never part of the product library.
#include "../sihl_b200/csrc/od_math.h"

extern "C" float shim_ciou_loss_row(const float *p, const float *t, float *grad)
{
    return sihl::ciou_loss_row(sihl::Box4{p[0], p[1], p[2], p[3]}, sihl::Box4{t[0], t[1], t[2], t[3]}, grad);
}

extern "C" void shim_ciou_matrix(const float *anchors, int na, const float *gt, int ng, float *out)
{
    for (int g = 0; g < ng; ++g) {
        sihl::BoxTerms tg = sihl::box_terms(sihl::Box4{gt[4 * g], gt[4 * g + 1], gt[4 * g + 2], gt[4 * g + 3]});
        for (int a = 0; a < na; ++a) {
            sihl::BoxTerms ta =
                sihl::box_terms(sihl::Box4{anchors[4 * a], anchors[4 * a + 1], anchors[4 * a + 2], anchors[4 * a + 3]});
            out[(long)a * ng + g] = sihl::ciou_pair(ta, tg);
        }
    }
}

extern "C" float shim_bce_logits(float x, float t) { return sihl::bce_logits(x, t); }
