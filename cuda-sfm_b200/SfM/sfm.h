// Source-compatible replacement of the reference's SfM/sfm.h: the same class,
// namespace, constructor and method names (SfM/sfm.h:20-60), so that the
// reference's driver (src/main.cpp:292-307, 337) recompiles unchanged against
// libsfmb200.  Private state is one opaque handle of the C ABI
// (include/sfmb200.h); every method forwards to it and then synchronises, which
// preserves the reference's observable behaviour (each of its stages ends with a
// device sync or a blocking copy).  Errors print and exit like the reference's
// checkCUDAError.
#pragma once

#include <cstdint>
#include <vector>

#include "../../include/sfmb200.h"
#include "common.h"

#if defined(__has_include)
#if __has_include(<CudaSift/cudaSift.h>)
#include <CudaSift/cudaSift.h>
#define SFMB200_HAVE_CUDASIFT 1
#endif
#endif
#ifndef SFMB200_HAVE_CUDASIFT
// Layout-compatible SiftPoint (CudaSift/cudaSift.h:6-22): 576 bytes, only
// xpos, ypos, match_xpos, match_ypos are read by fillXU.
typedef struct {
    float xpos, ypos, scale, sharpness, edgeness, orientation, score, ambiguity;
    int match;
    float match_xpos, match_ypos, match_error, subsampling;
    float empty[3];
    float data[128];
} SiftPoint;
#endif
static_assert(sizeof(SiftPoint) == 576, "SiftPoint layout");

#define checkCUDAErrorWithLine(msg) checkCUDAError(msg, __LINE__)

namespace kernels {
// singleton stopwatch (reference: kernels.h:26-30, declared at sfm.h:14-16)
inline Common::PerformanceTimer& timer() {
    static Common::PerformanceTimer t;
    return t;
}
}

namespace SfM {
#define cuda_block_size 256
class Image_pair {
    sfmb200_t* h_ = nullptr;
    int sampler_ = 0;      // explicit-argument estimateE: 0 independent rows, 1 disjoint permutation
    int image_count;
    int num_points;
    static void check(int rc, const char* what) {
        if (rc != 0) {
            fprintf(stderr, "sfmb200 error in %s: %s\n", what, sfmb200_last_error());
            exit(EXIT_FAILURE);
        }
    }

public:
    // k, k_inv: host 3x3 row-major; num_points = correspondences per pair.  image_count = 2 is the reference's
    // only use (main.cpp:298); image_count > 2 = a sequence: image_count - 1 consecutive pairs over index-aligned
    // tracks, filled with fillXU(d_pixels [pairs][n][4], n), every stage batched, chainViews() at the end.
    Image_pair(float k[9], float k_inv[9], int image_count, int num_points)
        : image_count(image_count), num_points(num_points) {
        if (image_count < 2 || image_count > 257) {
            fprintf(stderr, "Image_pair handles 2 to 257 images\n");
            exit(EXIT_FAILURE);
        }
        int max_h = num_points / 8 > 0 ? num_points / 8 : 1;
        check(sfmb200_create(k, k_inv, image_count - 1, num_points, max_h > 65536 ? max_h : 65536, &h_), "Image_pair");
    }
    ~Image_pair() { sfmb200_destroy(h_); }
    Image_pair(const Image_pair&) = delete;
    Image_pair& operator=(const Image_pair&) = delete;

    // ---- the reference's pipeline (same names, same order of calls) ----
    void fillXU(SiftPoint* data) {
        check(sfmb200_set_points_sift(h_, data, num_points), "fillXU");
        check(sfmb200_synchronize(h_), "fillXU");
    }
    // The reference draws H = N/8 disjoint samples from a std::shuffle seeded by
    // std::random_device (sfm.cu:95-104) and thresholds at 1e-6 (sfm.cu:220).
    // No arguments = exactly that: ONE permutation of the point indices cut into N/8 disjoint rows, drawn on the
    // device (SFMB200_OPT_SAMPLER = 1), threshold 1e-6.  The overload below takes H, seed and threshold explicitly
    // and samples every row independently unless setSampler(1) was called.
    void estimateE() {
        check(sfmb200_set_option(h_, SFMB200_OPT_SAMPLER, 1), "estimateE");
        estimateE(num_points / 8 > 0 ? num_points / 8 : 1, default_seed(), 1e-6f);
        check(sfmb200_set_option(h_, SFMB200_OPT_SAMPLER, sampler_), "estimateE");
    }
    void setSampler(int disjoint_permutation) {
        sampler_ = disjoint_permutation ? 1 : 0;
        check(sfmb200_set_option(h_, SFMB200_OPT_SAMPLER, sampler_), "setSampler");
    }
    void computePosecandidates() {
        check(sfmb200_pose_candidates(h_), "computePosecandidates");
        check(sfmb200_synchronize(h_), "computePosecandidates");
    }
    void choosePose() {
        check(sfmb200_choose_pose(h_), "choosePose");
        check(sfmb200_synchronize(h_), "choosePose");
    }
    void linear_triangulation() {
        check(sfmb200_triangulate(h_), "linear_triangulation");
        check(sfmb200_synchronize(h_), "linear_triangulation");
    }
    // device pointers, N x 4 floats each (positions (x, y, z, 1), colours all 1)
    void copyBoidsToVBO(float* vbodptr_positions, float* vbodptr_velocities) {
        check(sfmb200_copy_to_vbo(h_, 0, vbodptr_positions, vbodptr_velocities), "copyBoidsToVBO");
    }
    // same with a position scale and colours: mode 1 inlier / outlier, mode 2 depth ramp (see sfmb200.h)
    void copyBoidsToVBO(float* vbodptr_positions, float* vbodptr_colours, float scale, int mode, float z_near = 0.0f, float z_far = 1.0f) {
        check(sfmb200_copy_to_vbo_coloured(h_, 0, vbodptr_positions, vbodptr_colours, scale, mode, z_near, z_far), "copyBoidsToVBO");
    }
    // The reference's print-only self tests (sfm.cu:389-510), now asserting: each
    // runs its literal through the kernels.h-equivalent entry point and returns
    // whether the result matches the value the reference's comments expect.
    bool testSVD();
    bool testBatchedmult();
    bool testThrust_max();
    bool testInverse();
    bool testBatchedmultTranspose();
    bool testRow_extraction_kernel();
    bool testVecnorm();

    // ---- additive API (the reference keeps all results private) ----
    void estimateE(int H, uint64_t seed, float threshold, const int32_t* d_sample_rows = nullptr) {
        check(sfmb200_estimate_e(h_, d_sample_rows, H, seed, threshold), "estimateE");
        check(sfmb200_synchronize(h_), "estimateE");
    }
    void fillXU(const float* d_pixels_u1v1u2v2, int n) {
        check(sfmb200_set_points_xy(h_, d_pixels_u1v1u2v2, n), "fillXU");
        num_points = n;
    }
    // keep only matches with score > minScore && ambiguity < maxAmbiguity (CudaSift's FindHomography test);
    // returns the number kept
    int fillXU(SiftPoint* data, float minScore, float maxAmbiguity, int32_t* d_kept_index = nullptr) {
        int32_t kept = 0;
        check(sfmb200_set_points_sift_filtered(h_, data, num_points, minScore, maxAmbiguity, d_kept_index, &kept), "fillXU");
        return kept;
    }
    // ---- beyond the reference (its README.md:52,65-69 future work) ----
    // RANSAC that stops once log(1-confidence)/log(1-w^8) hypotheses were tried; returns the number tried
    int estimateEAdaptive(int maxH, uint64_t seed, float threshold, float confidence = 0.99f, int firstRound = 4096, int growth = 4) {
        int32_t used = 0;
        check(sfmb200_estimate_e_adaptive(h_, nullptr, maxH, firstRound, growth, seed, threshold, confidence, &used), "estimateEAdaptive");
        return used;
    }
    // LO-RANSAC refit of E on its inlier set; returns the accepted refits
    int refineE(int iterations = 4) {
        std::vector<int32_t> acc(pairs(), 0);
        check(sfmb200_refine_e(h_, iterations), "refineE");
        check(sfmb200_get_refit_iterations(h_, acc.data()), "refineE");
        return acc[0];
    }
    // CudaSift FindHomography semantics (matching.cu:907-1087) on this pair; returns the matches, H row-major 3x3
    int findHomography(float* H, int loops = 10000, float thresh = 5.0f, uint64_t seed = 0) {      // H [pairs()][9]
        std::vector<int32_t> matches(pairs(), 0);
        check(sfmb200_find_homography(h_, loops, seed, thresh, H, matches.data()), "findHomography");
        return matches[0];
    }
    // bundle adjustment of the chosen pose and the inlier points, with inlier re-selection; stats: see sfmb200.h
    void bundleAdjust(int outerRounds = 4, int iterations = 40, float* stats = nullptr)   /* stats [pairs()][8] */ {
        check(sfmb200_bundle_adjust(h_, outerRounds, iterations, stats), "bundleAdjust");
        check(sfmb200_synchronize(h_), "bundleAdjust");
    }
    // image_count > 2: consecutive pairs into one frame; any output pointer may be null (see sfmb200_chain_views)
    void chainViews(float* d_cloud_4xN, int32_t* d_seen, float* cameras_3x4, float* scales = nullptr, int32_t* links = nullptr) {
        check(sfmb200_chain_views(h_, d_cloud_4xN, d_seen, cameras_3x4, scales, links), "chainViews");
        check(sfmb200_synchronize(h_), "chainViews");
    }
    // global bundle adjustment of all cameras and points of the chained reconstruction (after chainViews with a cloud;
    // at most 17 images); cameras_3x4 / stats[8] may be null (see sfmb200_bundle_adjust_global)
    void bundleAdjustGlobal(float* d_cloud_4xN, const int32_t* d_seen, int iterations = 30, float* cameras_3x4 = nullptr, float* stats = nullptr) {
        check(sfmb200_bundle_adjust_global(h_, d_cloud_4xN, d_seen, iterations, cameras_3x4, stats), "bundleAdjustGlobal");
        check(sfmb200_synchronize(h_), "bundleAdjustGlobal");
    }
    int pairs() const { return image_count - 1; }
    void setCompat(bool reference_semantics) { check(sfmb200_set_option(h_, SFMB200_OPT_COMPAT, reference_semantics), "setCompat"); }
    // host getters; with image_count > 2 the arrays hold one entry per pair: E [pairs()][9], P [pairs()][64]
    void getE(float* E) { check(sfmb200_get_E(h_, E), "getE"); }
    void getPoses(float* P) { check(sfmb200_get_poses(h_, P), "getPoses"); }
    int getPoseIndex(int pair = 0) {
        std::vector<int32_t> i(pairs(), 0);
        check(sfmb200_get_pose_index(h_, i.data()), "getPoseIndex");
        return i[pair];
    }
    int getBest(int* inliers = nullptr, int pair = 0) {
        std::vector<int32_t> idx(pairs(), 0), cnt(pairs(), 0);
        check(sfmb200_get_best(h_, idx.data(), cnt.data()), "getBest");
        if (inliers) *inliers = cnt[pair];
        return idx[pair];
    }
    void getPoints(float* d_4xN) {
        check(sfmb200_get_points(h_, 0, d_4xN), "getPoints");
        check(sfmb200_synchronize(h_), "getPoints");
    }
    void getPointsHost(float* h_4xN) { check(sfmb200_get_points_host(h_, 0, h_4xN), "getPointsHost"); }
    void getInlierCounts(int32_t* d_counts) { check(sfmb200_get_inlier_counts(h_, 0, d_counts), "getInlierCounts"); }
    void getInlierMask(uint8_t* d_mask) { check(sfmb200_get_inlier_mask(h_, 0, d_mask), "getInlierMask"); }
    sfmb200_t* handle() { return h_; }

private:
    static uint64_t default_seed() {
        return (uint64_t)std::chrono::high_resolution_clock::now().time_since_epoch().count();
    }
};
}  // namespace SfM
