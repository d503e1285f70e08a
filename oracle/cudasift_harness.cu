// CudaSift front-end harness (TEST INFRASTRUCTURE ONLY - never linked into the
// product).  BASELINE config 1 is "data/dino image pair as the reference runs it
// (CudaSift matches)": the reference's src/main.cpp:251-282 loads two grey
// images, runs the vendored CudaSift ExtractSift on both and MatchSiftData, and
// hands siftData1.d_data to SfM::Image_pair::fillXU.  This file drives the
// UNMODIFIED CudaSift sources (compiled where they lie by oracle/Makefile:
// CudaSift/cudaImage.cu, cudaSiftH.cu, matching.cu) with main.cpp's own
// parameters so the correspondences of config 1 can be exported as a fixture
// (tests/golden/make_dino_cudasift_fixture.py) from a run on the GPU box.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include "CudaSift/cudaImage.h"
#include "CudaSift/cudaSift.h"

extern "C" {

// img1 / img2: w x h grey images as float (0..255), exactly what
// cv::imread(path, 0).convertTo(CV_32FC1) yields in main.cpp:251-252.
// out: [max_out][8] floats per feature of image 1:
//   xpos, ypos, match_xpos, match_ypos, score, ambiguity, match (index as float), match_error
// Returns the number of features of image 1 (siftData1.numPts), or < 0 on error;
// *n2 = features of image 2.  sift_out (optional): the raw SiftPoint array of
// image 1 (576 B records), what fillXU reads on the device.
int cudasift_match_pair(const float* img1, const float* img2, int w, int h, float* out, int max_out, int* n2,
                        void* sift_out) {
    InitCuda(0);                                                       // main.cpp:262
    CudaImage a, b;
    a.Allocate(w, h, iAlignUp(w, 128), false, NULL, (float*)img1);     // main.cpp:264-265
    b.Allocate(w, h, iAlignUp(w, 128), false, NULL, (float*)img2);
    a.Download();
    b.Download();
    SiftData s1, s2;
    float initBlur = 1.5f;                                             // main.cpp:270-271
    float thresh = 1.0f;
    InitSiftData(s1, 32768, true, true);
    InitSiftData(s2, 32768, true, true);
    float* tmp = AllocSiftTempMemory(w, h, 5, false);                  // main.cpp:277-280
    ExtractSift(s1, a, 5, initBlur, thresh, 0.0f, false, tmp);
    ExtractSift(s2, b, 5, initBlur, thresh, 0.0f, false, tmp);
    FreeSiftTempMemory(tmp);
    MatchSiftData(s1, s2);                                             // main.cpp:282
    cudaDeviceSynchronize();
    if (cudaGetLastError() != cudaSuccess) return -1;
    const int n = s1.numPts;
    if (n2) *n2 = s2.numPts;
    // the device array is what fillXU consumes; read it back whole
    SiftPoint* hp = new SiftPoint[n > 0 ? n : 1];
    cudaMemcpy(hp, s1.d_data, sizeof(SiftPoint) * (size_t)n, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n && i < max_out; i++) {
        float* o = out + 8 * (size_t)i;
        o[0] = hp[i].xpos; o[1] = hp[i].ypos; o[2] = hp[i].match_xpos; o[3] = hp[i].match_ypos;
        o[4] = hp[i].score; o[5] = hp[i].ambiguity; o[6] = (float)hp[i].match; o[7] = hp[i].match_error;
    }
    if (sift_out) memcpy(sift_out, hp, sizeof(SiftPoint) * (size_t)(n < max_out ? n : max_out));
    delete[] hp;
    FreeSiftData(s1);
    FreeSiftData(s2);
    return n;
}

}  // extern "C"
