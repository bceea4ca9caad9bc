// Tester sink on the GPU: what the reference does per batch on the host — `.argmax(dim=1).cpu().tolist()`
// then python lists of {task: int} dicts (probing/test/singletask_tester.py:95-98, base_tester.py:47-57) and, at the end,
// sklearn accuracy / confusion_matrix over them (datasets_vlm/evaluate_dataset.py:74-97,130-146) — kept on the device
// until the dataset is done: labels are appended to one [N, T] buffer, and the per-task confusion counts accumulate in
// integer atomics (order-independent, so bit-exact). One D2H copy per dataset instead of one sync per batch.
#include "common.cuh"

namespace vlmfw {
namespace {

// counts [T][C][C] (row = ground truth, column = prediction, as sklearn.metrics.confusion_matrix); gt < 0 = missing label
// (the evaluator skips those rows, evaluate_dataset.py:78); labels outside [0, C) are counted in `dropped`.
__global__ void __launch_bounds__(256) confusion_update_kernel(const int32_t* __restrict__ pred,
                                                               const int64_t* __restrict__ gt, int rows, int tasks,
                                                               int classes, int32_t* __restrict__ counts,
                                                               int32_t* __restrict__ dropped) {
    const int total = rows * tasks;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int t = i % tasks;
        const int64_t g = gt[i];
        const int p = pred[i];
        if (g < 0) continue;
        if (g >= classes || p < 0 || p >= classes) {
            atomicAdd(dropped, 1);
            continue;
        }
        atomicAdd(&counts[(static_cast<size_t>(t) * classes + static_cast<int>(g)) * classes + p], 1);
    }
}

}  // namespace
}  // namespace vlmfw

using namespace vlmfw;

extern "C" int vlmfw_confusion_update(const int32_t* pred, const int64_t* gt, int32_t rows, int32_t tasks,
                                      int32_t classes, int32_t* counts, int32_t* dropped, void* stream) {
    VLMFW_REQUIRE(pred && gt && counts && dropped, "confusion_update: null pointer");
    VLMFW_REQUIRE(rows >= 0 && tasks > 0 && classes > 0 && static_cast<int64_t>(rows) * tasks < (1ll << 31),
                  "confusion_update: bad shape rows=%d tasks=%d classes=%d", rows, tasks, classes);
    if (rows == 0) return VLMFW_OK;
    int rc = device_check();
    if (rc != VLMFW_OK) return rc;
    const int total = rows * tasks;
    int grid = (total + 255) / 256;
    if (grid > 4 * num_sms()) grid = 4 * num_sms();
    confusion_update_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(pred, gt, rows, tasks, classes, counts,
                                                                                 dropped);
    VLMFW_LAUNCH_CHECK();
    return VLMFW_OK;
}
