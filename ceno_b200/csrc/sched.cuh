// sched.cuh — memory-aware chip-proof lane scheduler, included by cabi.cu.
//
// Host-side counterpart of ChipScheduler::execute (reference ceno_zkvm/src/scheme/scheduler.rs:109-400,
// docs/src/concurrent-chip-proving.md): chip proofs are independent jobs; the reference runs them on 1..8 "lanes"
// (default 4), one OS thread + one non-default CUDA stream each, under a greedy backfilling policy:
//   1. sort the tasks by estimated memory, descending ("big rocks first");
//   2. while a lane is free, launch the FIRST pending task whose booking fits what is left of the budget
//      (skipping bigger ones — backfilling);
//   3. when nothing fits (or all lanes are busy) block until a running task completes and releases its booking;
//   4. if nothing fits and nothing is running, the remaining tasks can never run: report a deadlock error;
//   5. a lane is released only after everything the task submitted to the lane's stream has completed
//      (stream-local completion event); results come back ordered by task_id.
// The first failing task's status is returned after the in-flight tasks have drained (the reference's scope join).
// The per-task work itself (transcript fork, tower / main-sumcheck calls) is the caller's callback: it receives the
// lane's stream and passes it to the cg_* entry points, which are re-entrant per (ctx, stream).
#pragma once
#include <chrono>
#include <condition_variable>
#include <deque>
#include <thread>

struct SchedShared {
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<uint32_t> queue;        // indices (into the sorted order) handed to the lanes
    std::deque<uint32_t> done;         // sorted-order indices of completed tasks
    bool closing = false;
};

CG_EXPORT int cg_sched_execute(cg_ctx* ctx, const cg_sched_task* tasks, uint32_t n_tasks, uint32_t lanes, uint64_t mem_budget_bytes,
                               cg_sched_fn fn, void* user, cg_sched_result* results) {
    using clk = std::chrono::steady_clock;
    if (!fn || (n_tasks && (!tasks || !results))) return set_err(ctx, CG_ERR_INVALID, "cg_sched_execute: null argument");
    if (lanes == 0) lanes = CG_SCHED_DEFAULT_LANES;
    if (lanes > CG_SCHED_MAX_LANES)   // CENO_CHIP_PROVING_LANES accepts 1 through 8 (scheduler.rs:71-85)
        return set_err(ctx, CG_ERR_INVALID, "cg_sched_execute: lanes must be 1 through 8");
    if (n_tasks == 0) return CG_OK;
    if (mem_budget_bytes == 0) {
        if (!ctx) return set_err(ctx, CG_ERR_INVALID, "cg_sched_execute: a memory budget is required without a context");
        size_t fr = 0, tot = 0;
        cudaSetDevice(ctx->device);
        CU(ctx, cudaMemGetInfo(&fr, &tot));
        mem_budget_bytes = fr + ctx->reserved - ctx->used;   // free device memory + what the pool holds but does not use
    }
    // 1. big rocks first (stable: equal sizes keep the caller's order)
    std::vector<uint32_t> order(n_tasks);
    for (uint32_t i = 0; i < n_tasks; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return tasks[a].estimated_memory_bytes > tasks[b].estimated_memory_bytes; });
    auto booking = [&](uint32_t i) { return tasks[i].booked_memory_bytes ? tasks[i].booked_memory_bytes : tasks[i].estimated_memory_bytes; };

    const uint32_t n_lanes = std::min(lanes, n_tasks);
    std::vector<cudaStream_t> streams(n_lanes, nullptr);
    if (ctx) {
        cudaSetDevice(ctx->device);
        for (uint32_t l = 0; l < n_lanes; l++) {
            cudaError_t e = cudaStreamCreateWithFlags(&streams[l], cudaStreamNonBlocking);
            if (e != cudaSuccess) {
                for (uint32_t j = 0; j < l; j++) cudaStreamDestroy(streams[j]);
                return set_err(ctx, CG_ERR_CUDA, std::string("failed to acquire CUDA stream for lane ") + std::to_string(l) + ": " + cudaGetErrorString(e));
            }
        }
    }
    std::vector<cg_sched_result> res(n_tasks);   // indexed like `tasks`
    SchedShared sh;
    const auto t_start = clk::now();
    auto ms_since = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };

    std::vector<std::thread> workers;
    for (uint32_t lane = 0; lane < n_lanes; lane++) {
        workers.emplace_back([&, lane]() {
            if (ctx) cudaSetDevice(ctx->device);
            for (;;) {
                uint32_t i;
                {
                    std::unique_lock<std::mutex> lk(sh.mu);
                    sh.cv_work.wait(lk, [&] { return sh.closing || !sh.queue.empty(); });
                    if (sh.queue.empty()) return;
                    i = sh.queue.front();
                    sh.queue.pop_front();
                }
                cg_sched_result& r = res[i];
                const auto t0 = clk::now();
                r.queue_delay_ms = ms_since(t_start, t0);
                r.lane_id = lane;
                int st;
                try {   // a failing callback must not take the scheduler down (the reference catches worker panics)
                    st = fn(user, i, tasks[i].task_id, lane, (cg_stream)streams[lane]);
                } catch (...) {
                    st = CG_ERR_STATE;
                }
                const auto t1 = clk::now();
                r.host_execution_ms = ms_since(t0, t1);
                // neither the booking nor the lane is released before the lane's stream has drained
                if (ctx && streams[lane]) {
                    cudaError_t e = cudaStreamSynchronize(streams[lane]);
                    if (e != cudaSuccess && st == CG_OK) {
                        st = CG_ERR_CUDA;
                        set_err(ctx, CG_ERR_CUDA, std::string("CUDA completion event failed for task ") + std::to_string(tasks[i].task_id) + " on lane " +
                                                      std::to_string(lane) + ": " + cudaGetErrorString(e));
                    }
                }
                r.event_wait_ms = ms_since(t1, clk::now());
                r.status = st;
                {
                    std::lock_guard<std::mutex> g(sh.mu);
                    sh.done.push_back(i);
                }
                sh.cv_done.notify_one();
            }
        });
    }

    // 2-4. greedy backfilling
    std::vector<uint32_t> pending(order);
    uint64_t booked = 0;
    uint32_t inflight = 0, launch_seq = 0;
    int first_err = CG_OK;
    bool deadlock = false;
    auto handle_done = [&](uint32_t i) {
        booked -= booking(i);
        inflight--;
        if (res[i].status != CG_OK && first_err == CG_OK) first_err = res[i].status;
    };
    {
        std::unique_lock<std::mutex> lk(sh.mu);
        while ((!pending.empty() && first_err == CG_OK) || inflight > 0) {
            while (!sh.done.empty()) { handle_done(sh.done.front()); sh.done.pop_front(); }
            if (first_err != CG_OK && inflight == 0) break;
            if (first_err == CG_OK && inflight < n_lanes) {
                size_t pos = pending.size();
                for (size_t p = 0; p < pending.size(); p++)
                    if (booking(pending[p]) <= mem_budget_bytes - booked) { pos = p; break; }
                if (pos < pending.size()) {
                    const uint32_t i = pending[pos];
                    pending.erase(pending.begin() + pos);
                    booked += booking(i);
                    inflight++;
                    res[i].task_id = tasks[i].task_id;
                    res[i].launch_seq = launch_seq++;
                    res[i].booked_total_at_launch = booked;
                    sh.queue.push_back(i);
                    sh.cv_work.notify_one();
                    continue;
                }
            }
            if (inflight == 0) {
                if (pending.empty()) continue;
                deadlock = true;   // nothing runs and nothing fits: the rest can never be scheduled
                break;
            }
            sh.cv_done.wait(lk, [&] { return !sh.done.empty(); });
        }
        sh.closing = true;
    }
    sh.cv_work.notify_all();
    for (auto& w : workers) w.join();
    for (auto s : streams) if (s) cudaStreamDestroy(s);

    // 5. results ordered by task_id; tasks that never ran are marked CG_ERR_STATE
    for (uint32_t i : pending) { res[i].task_id = tasks[i].task_id; res[i].status = CG_ERR_STATE; res[i].lane_id = UINT32_MAX; res[i].launch_seq = UINT32_MAX; }
    std::vector<uint32_t> by_id(n_tasks);
    for (uint32_t i = 0; i < n_tasks; i++) by_id[i] = i;
    std::stable_sort(by_id.begin(), by_id.end(), [&](uint32_t a, uint32_t b) { return tasks[a].task_id < tasks[b].task_id; });
    for (uint32_t k = 0; k < n_tasks; k++) results[k] = res[by_id[k]];
    if (deadlock) {
        std::string msg = "Deadlock: Remaining tasks are too big for the memory pool: max=" + std::to_string(mem_budget_bytes) + " B, booked=" +
                          std::to_string(booked) + " B, pending=[";
        for (size_t p = 0; p < pending.size(); p++)
            msg += (p ? "; id=" : "id=") + std::to_string(tasks[pending[p]].task_id) + " booked=" + std::to_string(booking(pending[p])) + " B";
        return set_err(ctx, CG_ERR_OOM, msg + "]");
    }
    if (first_err != CG_OK && ctx) {
        std::lock_guard<std::mutex> g(ctx->mu);
        if (ctx->err.empty()) ctx->err = "cg_sched_execute: a task failed";
    }
    return first_err;
}

// Lane streams for callers that run their own threads (the reference binds one stream per OS thread,
// gkr_iop/src/gpu/mod.rs:79-154; get_pool_stream, scheduler.rs:437).
CG_EXPORT int cg_stream_create(cg_ctx* ctx, cg_stream* out) {
    if (!ctx || !out) return CG_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStream_t s = nullptr;
    CU(ctx, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *out = (cg_stream)s;
    return CG_OK;
}
CG_EXPORT int cg_stream_destroy(cg_ctx* ctx, cg_stream s) {
    if (!ctx || !s) return CG_ERR_INVALID;
    cudaSetDevice(ctx->device);
    CU(ctx, cudaStreamDestroy((cudaStream_t)s));
    return CG_OK;
}
