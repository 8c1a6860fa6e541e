// Library introspection entry points.
#include "common.cuh"
#include <atomic>
#include "build/source_hash.inc"     // written by tensorly_b200/build.py: sha256 of the sources of this build

namespace tlb200 {
int64_t launches();
static thread_local const char* g_last_path = "none";
void set_last_path(const char* name) { g_last_path = name; }
static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int64_t launches() { return g_launches.load(std::memory_order_relaxed); }
}  // namespace tlb200

extern "C" int tlb200_version(void) { return 100; }  // 0.1.0
extern "C" const char* tlb200_build_arch(void) { return "sm_100a"; }
extern "C" const char* tlb200_source_hash(void) { return TLB200_SOURCE_HASH; }
extern "C" int64_t tlb200_launch_count(void) { return tlb200::launches(); }
extern "C" const char* tlb200_last_path(void) { return tlb200::g_last_path; }
extern "C" const char* tlb200_status_string(int status) {
    switch (status) {
        case TLB200_OK: return "ok";
        case TLB200_EINVAL: return "invalid argument (shape / mode / rank / dtype / pointer)";
        case TLB200_EWORKSPACE: return "workspace too small";
        case TLB200_ECUDA: return "CUDA runtime error";
        case TLB200_EUNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}
