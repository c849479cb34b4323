// trace.cuh -- NVTX ranges around the entry points and, with SGLB200_TRACE=1, CUDA-event timers per hop printed to stderr.
// (The reference's only tracing on this path is time.time() around preprocess, sgl/tasks/node_classification.py:48-51;
// SURVEY.md section 5.)  NVTX3 is header-only and costs one predictable branch when no profiler is attached.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

namespace sglb200 {

struct TraceRange {
    explicit TraceRange(const char *name) { nvtxRangePushA(name); }
    ~TraceRange() { nvtxRangePop(); }
    TraceRange(const TraceRange &) = delete;
    TraceRange &operator=(const TraceRange &) = delete;
};

inline bool trace_enabled()
{
    static const bool on = [] {
        const char *e = getenv("SGLB200_TRACE");
        return e && atoi(e) != 0;
    }();
    return on;
}

// mark() before the first hop and after every hop; report() waits for the last event (tracing mode only -- the untraced
// path never synchronises) and prints one line per hop.
class HopTimer {
public:
    explicit HopTimer(const char *what) : what_(what), on_(trace_enabled()) {}
    ~HopTimer()
    {
        for (cudaEvent_t e : events_) cudaEventDestroy(e);
    }
    void mark(cudaStream_t stream)
    {
        if (!on_) return;
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, stream);
        events_.push_back(e);
    }
    void report()
    {
        if (!on_ || events_.size() < 2) return;
        if (cudaEventSynchronize(events_.back()) != cudaSuccess) return;
        float total = 0.0f;
        for (size_t k = 1; k < events_.size(); ++k) {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, events_[k - 1], events_[k]);
            fprintf(stderr, "sglb200 trace: %s hop %zu  %.3f ms\n", what_, k, ms);
            total += ms;
        }
        fprintf(stderr, "sglb200 trace: %s %zu hops  %.3f ms\n", what_, events_.size() - 1, total);
    }

private:
    const char *what_;
    bool on_;
    std::vector<cudaEvent_t> events_;
};

}  // namespace sglb200
