/*
 * ppf_oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker / the timed host baseline.
 *
 * PARITY UNPINNED: the arithmetic restated here lives in PCL (features/src/pfh.cpp,
 * features/impl/ppf.hpp, registration/impl/ppf_registration.hpp,
 * registration/src/ppf_registration.cpp), which is neither vendored in the reference
 * nor installable in the build container, and the reference ships no tests or golden
 * vectors for the path.  The restatement follows SURVEY.md Appendix A.
 *
 * All clouds are row-major float32 N×6 = [x y z nx ny nz], the layout the reference
 * hands to its PPF engine (reference include/CloudProcessing.h:163-190).
 */
#ifndef PPF_ORACLE_H
#define PPF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* feature functor: what PCL executes, and the two Drost alternates */
enum { ORACLE_FEATURE_PCL_PFH = 0, ORACLE_FEATURE_DROST_COS = 1, ORACLE_FEATURE_DROST_ANGLE = 2 };
/* alpha binning: mode A = PCL >= 1.12 (canonical), mode B = PCL 1.8-1.11 */
enum { ORACLE_ALPHA_MODE_A = 0, ORACLE_ALPHA_MODE_B = 1 };
/* columns of the voting accumulator (SURVEY.md A.4 `aux_size`): CEIL = ceil(2*pi/step), current PCL and the
 * default; FLOOR_DROP = floor(2*pi/step) with the votes for the missing last bin lost, PCL 1.8 .. 1.11;
 * FLOOR_CLAMP = floor with those votes added to the last column (round 1 of this repository). */
enum { ORACLE_NALPHA_CEIL = 0, ORACLE_NALPHA_FLOOR_DROP = 1, ORACLE_NALPHA_FLOOR_CLAMP = 2 };

/* A.1 / A.1' : one pair feature.  Returns 1 when the pair is valid. */
int oracle_pair_feature(int feature_mode, const float *p1, const float *n1, const float *p2,
                        const float *n2, float *f /*[4]*/);

/* A.2 : planar angle of point p in the frame that moves (p_r, n_r) to the origin / x-axis. */
float oracle_alpha(const float *p_r, const float *n_r, const float *p);

/* A.2 : rigid frame of a reference point; R row-major [9], t [3]. */
void oracle_ref_frame(const float *p_r, const float *n_r, float *R, float *t);

/* A.2 : PPFEstimation::computeFeature over all ordered pairs.  out is N*N*5 floats
 * {f1,f2,f3,f4,alpha_m}, row-major [i*N+j]; invalid pairs (i==j or failure) are all-NaN.
 * Returns the number of valid pairs. */
size_t oracle_ppf_estimation(int feature_mode, const float *cloud, size_t n, float *out);
size_t oracle_ppf_estimation_mt(int feature_mode, const float *cloud, size_t n, float *out, int n_threads);

/* A.3 : PPFHashMapSearch */
typedef struct oracle_hashmap oracle_hashmap;
oracle_hashmap *oracle_hashmap_create(float angle_step, float dist_step);
void oracle_hashmap_destroy(oracle_hashmap *hm);
/* setInputFeatureCloud: feats = count*5 floats, count = n*n. NaN pairs are skipped. */
void oracle_hashmap_set_features(oracle_hashmap *hm, const float *feats, size_t count);
/* the same container built by n_threads threads (sharded by key; identical bucket contents) — keeps the CPU side of
 * the full-size parity tests and of bench.py within minutes; n_threads <= 1 is the single container */
void oracle_hashmap_set_features_mt(oracle_hashmap *hm, const float *feats, size_t count, int n_threads);
float oracle_hashmap_model_diameter(const oracle_hashmap *hm);
size_t oracle_hashmap_num_entries(const oracle_hashmap *hm);
size_t oracle_hashmap_num_keys(const oracle_hashmap *hm);
/* quantisation used by the map: d[4] = floor(f/step) */
void oracle_hashmap_quantise(const oracle_hashmap *hm, const float *f, int32_t *d);
/* nearestNeighborSearch: writes up to cap (i,j) pairs, canonical order (i asc, j asc).
 * Returns the full bucket length. */
size_t oracle_hashmap_query(const oracle_hashmap *hm, float f1, float f2, float f3, float f4,
                            uint64_t *pairs /*[cap][2]*/, size_t cap);
size_t oracle_hashmap_query_key(const oracle_hashmap *hm, const int32_t *d, uint64_t *pairs,
                                size_t cap);
/* dump every distinct key (4 ints each) and its bucket length; arrays sized num_keys */
void oracle_hashmap_dump_keys(const oracle_hashmap *hm, int32_t *keys, uint32_t *lengths);

/* the column rule of every voting call on this map (default ORACLE_NALPHA_CEIL) */
void oracle_hashmap_set_nalpha_rule(oracle_hashmap *hm, int nalpha_rule);
/* A.4 : number of alpha bins */
uint32_t oracle_num_alpha_bins(float angle_step, int nalpha_rule);
/* A.4 : bin of one (alpha_m, alpha_s) pair. Returns UINT32_MAX for NaN, UINT32_MAX - 1 for a vote the rule drops. */
uint32_t oracle_alpha_bin(int alpha_mode, int nalpha_rule, float angle_step, float alpha_m, float alpha_s);

/* A.4 : per-scene-pair quantities of one reference point (debug / parity):
 * for every scene point s writes in_radius[s] (0/1; 0 for s==s_r and for failed pairs),
 * d[s][4] and alpha_s[s].  Returns the number of in-radius valid pairs. */
size_t oracle_scene_pairs(const oracle_hashmap *hm, int feature_mode, const float *scene,
                          size_t n_s, size_t s_r, uint8_t *in_radius, int32_t *d,
                          float *alpha_s);

/* A.4 : the accumulator of one reference point, acc is n_m * n_alpha uint32, zeroed here.
 * Returns total votes cast. */
uint64_t oracle_vote_accumulate(const oracle_hashmap *hm, int feature_mode, int alpha_mode,
                                size_t n_m, const float *scene, size_t n_s, size_t s_r,
                                uint32_t *acc);
/* integer half only: accumulate from externally supplied per-pair keys and alpha_s
 * (used to check the device's index work bit-exactly given the device's own float results) */
uint64_t oracle_vote_accumulate_from_pairs(const oracle_hashmap *hm, int alpha_mode, size_t n_m,
                                           size_t n_pairs, const int32_t *d /*[n][4]*/,
                                           const float *alpha_s, uint32_t *acc);

uint64_t oracle_vote_accumulate_from_pairs_mt(const oracle_hashmap *hm, int alpha_mode, size_t n_m, size_t n_pairs,
                                              const int32_t *d, const float *alpha_s, uint32_t *acc, int n_threads);

/* one hypothesis per reference point (64 bytes, same record the device emits) */
typedef struct oracle_hypothesis {
    float pose[12];  /* 3x4 row-major, model -> scene */
    uint32_t votes;
    uint32_t model_index;  /* i*  */
    uint32_t alpha_bin;    /* j*  */
    uint32_t scene_index;  /* s_r */
} oracle_hypothesis;

/* A.4 : voting loop over reference points ref_first, ref_first+ref_step, ... (ref_count of
 * them).  n_threads <= 1 is PCL as shipped; > 1 uses OpenMP over reference points with
 * thread-private accumulators.  stats (optional, [4]) receives pairs examined, pairs in
 * radius, non-empty lookups, votes. */
int oracle_vote(const oracle_hashmap *hm, int feature_mode, int alpha_mode, const float *model,
                size_t n_m, const float *scene, size_t n_s, size_t ref_first, size_t ref_step,
                size_t ref_count, int n_threads, oracle_hypothesis *hyps, uint64_t *stats);

/* the same loop over an explicit list of reference points (bench.py's fixed CPU sample) */
int oracle_vote_refs(const oracle_hashmap *hm, int feature_mode, int alpha_mode, const float *model, size_t n_m,
                     const float *scene, size_t n_s, const size_t *refs, size_t ref_count, int n_threads,
                     oracle_hypothesis *hyps, uint64_t *stats);

/* A.4 : pose of one peak (model_index, alpha_bin) for scene reference s_r */
void oracle_peak_pose(int alpha_mode, float angle_step, const float *model, size_t model_index,
                      uint32_t alpha_bin, const float *scene, size_t s_r, float *pose12);

/* A.5 : clusterPoses.  out_poses [3][16] row-major 4x4, out_votes [3]; returns #results.
 * assignment (optional, [n]) receives the cluster creation index of every input hypothesis
 * (in input order), n_clusters (optional) the number of clusters. */
size_t oracle_cluster(const oracle_hypothesis *hyps, size_t n, float pos_thr, float rot_thr,
                      float *out_poses, uint32_t *out_votes, uint32_t *assignment,
                      size_t *n_clusters);

/* A.5 : posesWithinErrorBounds on two 3x4 poses */
int oracle_poses_within(const float *a12, const float *b12, float pos_thr, float rot_thr);

/* A.4 tail : transformPointCloud (xyz only) ; out is n*3 floats */
void oracle_transform(const float *cloud, size_t n, const float *pose16, float *out_xyz);

/* PPFRegistration::align in one call. final16 = getFinalTransformation(); returns #results */
size_t oracle_register(const oracle_hashmap *hm, int feature_mode, int alpha_mode,
                       const float *model, size_t n_m, const float *scene, size_t n_s,
                       size_t ref_rate, float pos_thr, float rot_thr, int n_threads,
                       float *final16, float *out_poses, uint32_t *out_votes, uint64_t *stats);

int oracle_max_threads(void);

/* "Next" row of the path: the ICP refinement the reference runs on the top PPF
 * poses — reference include/CloudProcessing.h:465-470 / :518-523, ICP(100, 0.005f, 2.5f, 8) and
 * registerModelToScene(model, scene, poses).  Restated in icp_oracle.cpp from opencv_contrib
 * surface_matching/src/icp.cpp as recalled (parity unpinned).  poses16: n_poses row-major 4x4 doubles,
 * model -> scene, refined in place; residuals[n_poses] and the total iteration count are optional. */
int oracle_icp_refine(const float *model, size_t n_model, const float *scene, size_t n_scene, int max_iterations,
                      float tolerance, float rejection_scale, int num_levels, double *poses16, size_t n_poses,
                      double *residuals, uint64_t *iterations_run);

/* "Next" row: the scene pre-processing between the YOLO crop and the PPF engine —
 * reference include/CloudProcessing.h:359-377 (pcl::VoxelGrid), :340-358 (pcl::StatisticalOutlierRemoval),
 * :378-401 (pcl::NormalEstimationOMP, k = 30), :402-427 (curvature edges), :163-190 (normals re-normalised).
 * Restated in prep_oracle.cpp from the PCL files named there (parity unpinned).  xyz clouds are float32 rows
 * of `stride` floats starting with x y z. */
size_t oracle_voxel_grid(const float *xyz, size_t n, size_t stride, const float *leaf3, float *out_xyz, int *status);
void oracle_knn(const float *xyz, size_t n, size_t stride, int k, uint32_t *idx, float *d2, int n_threads);
size_t oracle_sor(const float *xyz, size_t n, size_t stride, int mean_k, double std_mul, float *distances, uint8_t *keep,
                  double *threshold, int n_threads);
void oracle_normals(const float *xyz, size_t n, size_t stride, int k, const float *viewpoint3, int cov_mode, float *out4,
                    int n_threads);
void oracle_renormalize_normals(float *nrm, size_t n, size_t stride);
/* SceneCropping (reference include/CloudProcessing.h:263-339, include/Camera.h:50-61) */
void oracle_frustum_corners(const float *depth, int rows, int cols, int bx, int by, int bw, int bh, double fx, double fy,
                            double ppx, double ppy, float *corners12);
size_t oracle_crop_pyramid(const float *xyz, size_t n, size_t stride, const float *corners12, uint8_t *keep);

/* The engine the reference actually calls — cv::ppf_match_3d::PPF3DDetector (opencv_contrib surface_matching),
 * reference include/CloudProcessing.h:205-236 (construction, trainModel), :442 (match) — restated in cvppf_oracle.cpp
 * as the checker of a row that is not built on the device yet.  Parity unpinned; see the
 * header of that file for what is known to be soft.  Clouds are N x 6 float32 [x y z nx ny nz]. */
typedef struct oracle_cv_detector oracle_cv_detector;
oracle_cv_detector *oracle_cv_create(double relative_sampling_step, double relative_distance_step, double num_angles);
void oracle_cv_destroy(oracle_cv_detector *d);
void oracle_cv_set_search_params(oracle_cv_detector *d, double position_threshold, double rotation_threshold);
size_t oracle_cv_sample(const float *pc, size_t n, float sample_step, float *out);
uint32_t oracle_cv_murmur(const void *key, int len, uint32_t seed);
uint32_t oracle_cv_pair(const float *p1n1, const float *p2n2, double angle_step, double distance_step, double *f);
size_t oracle_cv_train(oracle_cv_detector *d, const float *model, size_t n);
size_t oracle_cv_model_points(const oracle_cv_detector *d, float *out6);
size_t oracle_cv_match(const oracle_cv_detector *d, const float *scene, size_t n, double relative_scene_sample_step,
                       double relative_scene_distance, double *out_poses, uint32_t *out_votes, size_t cap, uint32_t *raw3,
                       size_t *n_refs, int n_threads);

/* the fork-only match_S2B as inferred (see cvppf_oracle.cpp): reference points from the scene, paired with the edge cloud */
size_t oracle_cv_match_s2b(const oracle_cv_detector *d, const float *scene, size_t n, const float *edge, size_t n_edge,
                           double relative_scene_sample_step, double relative_scene_distance, double *out_poses,
                           uint32_t *out_votes, size_t cap, uint32_t *raw3, size_t *n_refs, int n_threads);
/* parity hooks for the device engine: one bucket of the chained table as (i, ppfInd) pairs sorted by ppfInd (returns its
 * length), the table size, and the accumulator of one reference point of match / match_S2B (m * num_angles words) */
size_t oracle_cv_table_size(const oracle_cv_detector *d);
size_t oracle_cv_bucket(const oracle_cv_detector *d, size_t bucket, uint32_t *ppf_ind, size_t cap);
size_t oracle_cv_accumulator(const oracle_cv_detector *d, const float *scene, size_t n, const float *edge, size_t n_edge,
                             double relative_scene_distance, size_t sampled_index, uint32_t *acc);

#ifdef __cplusplus
}
#endif
#endif
