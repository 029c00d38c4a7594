// TEST DOUBLE, tests/ only: the five C-ABI entry points the OpenCV-shaped ICP shim calls, implemented over the CPU
// checker (oracle/icp_oracle.cpp) so that the shim's marshalling — cv::Mat rows -> cloud, Pose3D -> 16 doubles and
// back, residuals, parameters — can be checked where there is no GPU (tests/test_host.py).  Never linked into the
// product: libb200ppf.so has no CPU path.
#include <cstring>
#include <vector>

#include "../../include/b200ppf.h"
#include "../../oracle/ppf_oracle.h"

struct b200ppf_ctx {
    int unused;
};
struct b200ppf_cloud {
    std::vector<float> rows;  // N x 6
};

extern "C" {

int b200ppf_create(int, b200ppf_ctx **out) {
    *out = new b200ppf_ctx();
    return B200PPF_OK;
}
const char *b200ppf_last_error(const b200ppf_ctx *) { return "mock"; }
int b200ppf_cloud_upload(b200ppf_ctx *, const float *host, size_t n, size_t stride, size_t noff, b200ppf_cloud **out) {
    b200ppf_cloud *c = new b200ppf_cloud();
    c->rows.resize(n * 6);
    for (size_t i = 0; i < n; ++i) {
        std::memcpy(&c->rows[6 * i], host + i * stride, 3 * sizeof(float));
        std::memcpy(&c->rows[6 * i + 3], host + i * stride + noff, 3 * sizeof(float));
    }
    *out = c;
    return B200PPF_OK;
}
void b200ppf_cloud_free(b200ppf_cloud *c) { delete c; }
int b200ppf_icp_refine(b200ppf_ctx *, const b200ppf_cloud *model, const b200ppf_cloud *scene, const b200ppf_icp_params *p,
                       double *poses16, size_t n_poses, double *residuals, uint64_t *iterations) {
    return oracle_icp_refine(model->rows.data(), model->rows.size() / 6, scene->rows.data(), scene->rows.size() / 6,
                             p->max_iterations, p->tolerance, p->rejection_scale, p->num_levels, poses16, n_poses, residuals,
                             iterations);
}

}  // extern "C"
