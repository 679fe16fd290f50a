// TEST INFRASTRUCTURE ONLY -- scheduler of the SIMT emulator (see simt_emu.h).
#include "simt_emu.h"

#include <atomic>
#include <thread>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

static void simt_segv(int sig) {
    void* bt[64];
    int n = backtrace(bt, 64);
    const char msg[] = "simt_emu: fatal signal, backtrace:\n";
    (void)!write(2, msg, sizeof(msg) - 1);
    backtrace_symbols_fd(bt, n, 2);
    _exit(128 + sig);
}
__attribute__((constructor)) static void simt_install_handler() {
    if (getenv("LB_EMU_BACKTRACE")) {
        signal(SIGSEGV, simt_segv);
        signal(SIGABRT, simt_segv);
    }
}

namespace simt {

thread_local Ctx* cur = nullptr;
thread_local void* sched_sp = nullptr;
thread_local unsigned cur_site = 0;
bool capture_bt = getenv("LB_EMU_BT") != nullptr;
int capture_backtrace(void** out, int n) { return backtrace(out, n); }
void divergence_abort(unsigned a, unsigned b, unsigned lane_a, unsigned lane_b) {
    if (capture_bt && cur && cur->warp) {
        std::fprintf(stderr, "--- stack of lane %u\n", lane_a);
        backtrace_symbols_fd(cur->warp->bt[lane_a], cur->warp->bt_n[lane_a], 2);
        std::fprintf(stderr, "--- stack of lane %u\n", lane_b);
        backtrace_symbols_fd(cur->warp->bt[lane_b], cur->warp->bt_n[lane_b], 2);
    }
    std::fprintf(stderr, "simt_emu: WARP DIVERGENCE at a collective: lane %u is at line %u, lane %u at line %u "
                 "(site = line + 100000*(len(file)%%1000))\n", lane_a, a % 100000u, lane_b, b % 100000u);
    abort();
}

// x86-64 SysV context switch: saves callee-saved registers on the current stack, swaps stack pointers.
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_switch,.-simt_switch
)");

static const size_t STACK_BYTES = 256 * 1024;

struct CtaRun {
    const std::function<void()>* body;
    std::vector<Fiber> fibers;
    std::vector<Ctx> ctxs;
    std::vector<WarpSync> warps;
    CtaSync cta;
};
static thread_local CtaRun* running = nullptr;
static thread_local Fiber* starting = nullptr;

static void fiber_main() {
    Fiber* f = starting;
    Ctx* c = cur;
    (*running->body)();
    f->done = true;
    // leave the rendezvous populations
    c->warp->alive &= ~(1u << f->lane);
    c->cta->alive--;
    c->progress++;
    // a departing lane may complete a pending rendezvous
    WarpSync& w = *c->warp;
    if (w.arrived && (w.arrived & w.alive) == w.alive) {
        w.part[w.gen & 1] = w.arrived;
        w.arrived = 0;
        w.gen++;
    }
    CtaSync& cs = *c->cta;
    if (cs.alive && cs.arrived >= cs.alive) {
        cs.arrived = 0;
        cs.gen++;
    }
    void* dummy;
    simt_switch(&dummy, sched_sp);
    abort();
}

void yield() {
    Fiber* f = cur->fiber;
    Ctx* me = cur;
    simt_switch(&f->sp, sched_sp);
    cur = me;
}

static void run_cta(const std::function<void()>& body, dim3 grid, dim3 block, size_t smem, unsigned bx,
                    unsigned by, unsigned bz, std::vector<char*>& stacks) {
    unsigned nthreads = block.x * block.y * block.z;
    CtaRun run;
    run.body = &body;
    run.fibers.resize(nthreads);
    run.ctxs.resize(nthreads);
    run.warps.resize((nthreads + 31) / 32);
    run.cta.alive = nthreads;
    std::vector<char> dyn(smem ? smem : 1);
    while (stacks.size() < nthreads) stacks.push_back((char*)std::malloc(STACK_BYTES));
    for (unsigned t = 0; t < nthreads; t++) {
        Fiber& f = run.fibers[t];
        f.tid = t;
        f.lane = t & 31;
        f.warp = t >> 5;
        f.stack = stacks[t];
        run.warps[f.warp].alive |= 1u << f.lane;
        Ctx& c = run.ctxs[t];
        c.tid = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
        c.bid = {bx, by, bz};
        c.bdim = block;
        c.gdim = grid;
        c.fiber = &f;
        c.warp = &run.warps[f.warp];
        c.cta = &run.cta;
        c.dyn_smem = dyn.data();
        // initial stack: 6 callee-saved slots + return address (fiber_main); rsp%16==8 at entry
        uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;               // alignment pad (fake return address of fiber_main)
        *--sp = (void*)&fiber_main;    // popped by `ret`
        for (int i = 0; i < 6; i++) *--sp = nullptr;
        f.sp = sp;
    }
    running = &run;
    unsigned remaining = nthreads;
    uint64_t last_progress = 0;
    int stuck_rounds = 0;
    bool started_all = false;
    std::vector<char> started(nthreads, 0);
    while (remaining) {
        uint64_t prog = 0;
        for (unsigned t = 0; t < nthreads; t++) {
            Fiber& f = run.fibers[t];
            if (f.done) continue;
            cur = &run.ctxs[t];
            if (!started[t]) {
                started[t] = 1;
                starting = &f;
            }
            simt_switch(&sched_sp, f.sp);
            if (f.done) remaining--;
        }
        for (unsigned t = 0; t < nthreads; t++) prog += run.ctxs[t].progress;
        (void)started_all;
        if (prog == last_progress) {
            if (++stuck_rounds > 4) {
                std::fprintf(stderr, "simt_emu: deadlock in CTA (%u,%u,%u): %u threads blocked\n", bx, by, bz, remaining);
                abort();
            }
        } else
            stuck_rounds = 0;
        last_progress = prog;
    }
    running = nullptr;
    cur = nullptr;
}

static int emu_threads() {
    const char* e = std::getenv("LB_EMU_THREADS");
    int n = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
    return n < 1 ? 1 : n;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    uint64_t total = (uint64_t)grid.x * grid.y * grid.z;
    if (total == 0) return;
    std::atomic<uint64_t> next(0);
    auto worker = [&]() {
        std::vector<char*> stacks;
        while (true) {
            uint64_t i = next.fetch_add(1);
            if (i >= total) break;
            unsigned bx = (unsigned)(i % grid.x), by = (unsigned)((i / grid.x) % grid.y),
                     bz = (unsigned)(i / ((uint64_t)grid.x * grid.y));
            run_cta(body, grid, block, smem, bx, by, bz, stacks);
        }
        for (char* s : stacks) std::free(s);
    };
    int nt = emu_threads();
    if ((uint64_t)nt > total) nt = (int)total;
    std::vector<std::thread> ts;
    for (int t = 1; t < nt; t++) ts.emplace_back(worker);
    worker();
    for (auto& t : ts) t.join();
}

}  // namespace simt


// ---------------------------------------------------------------- guarded allocations (LB_EMU_GUARD)
#include <sys/mman.h>
#include <unistd.h>
#include <map>
#include <mutex>
namespace {
std::mutex g_guard_mu;
std::map<void*, std::pair<void*, size_t>> g_guard;   // user pointer -> (mapping, mapped bytes)
}
bool simt_guard_enabled() {
    static int on = getenv("LB_EMU_GUARD") ? 1 : 0;
    return on != 0;
}
cudaError_t simt_guard_malloc(void** p, size_t n) {
    size_t page = (size_t)sysconf(_SC_PAGESIZE);
    size_t body = ((n ? n : 1) + 255) & ~(size_t)255;          // the engine rounds to 256 anyway
    size_t pages = (body + page - 1) / page;
    size_t total = (pages + 2) * page;
    char* m = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) return 2;
    mprotect(m, page, PROT_NONE);
    mprotect(m + (pages + 1) * page, page, PROT_NONE);
    char* user = m + (pages + 1) * page - body;                  // the end of the buffer touches the guard page
    std::lock_guard<std::mutex> g(g_guard_mu);
    g_guard[user] = {m, total};
    *p = user;
    return 0;
}
cudaError_t simt_guard_free(void* p) {
    if (!p) return 0;
    std::lock_guard<std::mutex> g(g_guard_mu);
    auto it = g_guard.find(p);
    if (it == g_guard.end()) { std::free(p); return 0; }
    munmap(it->second.first, it->second.second);
    g_guard.erase(it);
    return 0;
}
