/* The per-box body of the reference's main() (src/YOLO_cropping_ppf_test.cpp:91-122) through the C ABI, as INTEGRATION.md
 * section 3c shows it: the frame uploaded once, one b200ppf_match_object call per YOLO box.  Compiled as C99 by
 * tests/test_host.py::test_header_is_plain_c (the header must stay free of C++). */
#include <stddef.h>
#include "b200ppf.h"
static void use(const double *pose, double residual, unsigned votes) { (void)pose; (void)residual; (void)votes; }
int run(b200ppf_ctx *ctx, const float *scene_xyz4, size_t n, const float *depth, int rows, int cols, const int (*boxes)[4], size_t n_boxes,
        double fx, double fy, double ppx, double ppy, const b200ppf_cloud *model, const b200ppf_table *table, float leaf, double outlier_thresh) {
    b200ppf_cloud *frame;
    if (b200ppf_cloud_upload_xyz(ctx, scene_xyz4, n, 4, &frame)) return -1;
    for (size_t i = 0; i < n_boxes; ++i) {
        float corners[12];
        b200ppf_frustum_corners(depth, rows, cols, boxes[i][0], boxes[i][1], boxes[i][2], boxes[i][3], fx, fy, ppx, ppy, corners);
        b200ppf_object_params prm;
        b200ppf_object_params_default(&prm);
        prm.leaf = leaf;
        prm.sor_stddev_mul = outlier_thresh;
        b200ppf_object_result r;
        if (b200ppf_match_object(ctx, frame, corners, model, table, &prm, &r, NULL, NULL) == B200PPF_OK) use(r.pose, r.residual, r.votes);
    }
    b200ppf_cloud_free(frame);
    return 0;
}
