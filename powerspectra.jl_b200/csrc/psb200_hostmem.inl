// psb200_hostmem.inl -- page-locked result buffers placed across the NUMA nodes of the host (included at the end of
// psb200.cu).
//
// Why: the result matrices of this path (N^2 x 8 B each: 302 MB at lmax 6143, six per benchmark step) leave the GPUs by
// DMA writes into ONE host array.  On a two-socket 8-GPU box that array sits on the socket of the thread that touched
// it first, and the four GPUs of the other socket write across the socket link: in the round-2 traces of the 8-GPU host
// call (profiles/r02_trace_n8_tt.log) GPUs 0-3 deliver their 38 MB in 1.1-1.7 ms, GPUs 4-7 need 2.7-3.0 ms.  An array
// whose 2 MB pieces alternate between the nodes sends half of every GPU's bytes to its own socket and loads the link
// equally in both directions.
//
// How: mmap + (mbind MPOL_INTERLEAVE, or -- where the container forbids the call -- first touch by threads bound to
// the CPUs of each node) + cudaHostRegister(portable).  Pages are locked by the registration, so automatic NUMA
// balancing cannot move them afterwards.  Nothing here is on a compute path.
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

namespace {

struct HostBlock { size_t bytes; int registered; };
std::mutex g_host_mutex;
std::map<void*, HostBlock> g_host_blocks;

constexpr size_t kPiece = size_t(2) << 20;      // placement granularity (one transparent huge page)

// CPUs of every NUMA node this process may run on: /sys/devices/system/node/node<k>/cpulist cut by the affinity mask
std::vector<std::pair<int, std::vector<int>>> numa_nodes_with_cpus()
{
    std::vector<std::pair<int, std::vector<int>>> out;
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return out;
    for (int node = 0; node < 64; ++node) {
        char path[96];
        snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
        FILE* f = fopen(path, "r");
        if (!f) continue;
        char line[4096] = {0};
        const bool got = fgets(line, sizeof line, f) != nullptr;
        fclose(f);
        if (!got) continue;
        std::vector<int> cpus;
        char* save = nullptr;
        for (char* tok = strtok_r(line, ",\n", &save); tok; tok = strtok_r(nullptr, ",\n", &save)) {
            int a = 0, b = 0;
            const int n = sscanf(tok, "%d-%d", &a, &b);
            if (n == 1) b = a;
            if (n < 1) continue;
            for (int c = a; c <= b && c < CPU_SETSIZE; ++c)
                if (CPU_ISSET(c, &allowed)) cpus.push_back(c);
        }
        if (!cpus.empty()) out.emplace_back(node, std::move(cpus));
    }
    return out;
}

// node of the page that holds p, or -1 (get_mempolicy may be filtered inside a container)
int page_node(const void* p)
{
#ifdef SYS_get_mempolicy
    int node = -1;
    const long rc = syscall(SYS_get_mempolicy, &node, nullptr, 0UL, (void*)p, 3UL /* MPOL_F_NODE | MPOL_F_ADDR */);
    return rc == 0 ? node : -1;
#else
    (void)p;
    return -1;
#endif
}

// Touch [base, base + bytes): piece k (2 MB) by a thread bound to node (k mod nnodes) when `by_node`, else by nthreads
// unbound threads (the policy set with mbind places the pages).
void touch_pieces(char* base, size_t bytes, const std::vector<std::pair<int, std::vector<int>>>& nodes, bool by_node)
{
    const size_t npieces = (bytes + kPiece - 1) / kPiece;
    const int nn = by_node ? (int)nodes.size() : 1;
    const int per = by_node ? 2 : 4;                        // threads per group
    std::vector<std::thread> th;
    for (int g = 0; g < nn; ++g)
        for (int t = 0; t < per; ++t)
            th.emplace_back([=, &nodes] {
                if (by_node) {
                    cpu_set_t set;
                    CPU_ZERO(&set);
                    for (int c : nodes[g].second) CPU_SET(c, &set);
                    sched_setaffinity(0, sizeof set, &set);  // this thread only (tid 0 = caller)
                }
                size_t mine = 0;
                for (size_t k = g; k < npieces; k += nn, ++mine) {
                    if ((int)(mine % per) != t) continue;
                    const size_t lo = k * kPiece, hi = std::min(bytes, lo + kPiece);
                    memset(base + lo, 0, hi - lo);
                }
            });
    for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

/* Page-locked host memory for result matrices.  policy 0: pages where the calling thread runs (what cudaHostAlloc
 * gives); policy 1: 2 MB pieces alternate between the NUMA nodes this process may run on (falls back to 0 on a
 * one-node host).  Returns NULL on failure (psb200_last_error).  Registration with CUDA is skipped, not failed, when no
 * device is visible, so the placement logic is testable on a CPU-only host. */
void* psb200_host_alloc(size_t bytes, int policy)
{
    if (bytes == 0 || policy < 0 || policy > 1) { fail(ERR_ARG, "host_alloc: bytes must be > 0 and policy 0 or 1"); return nullptr; }
    const size_t len = (bytes + kPiece - 1) / kPiece * kPiece;
    void* p = mmap(nullptr, len + kPiece, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) { fail(ERR_OOM, "host_alloc: mmap of %zu bytes failed", len); return nullptr; }
    // 2 MB-aligned start inside the mapping (the unused head and tail are returned)
    char* base = (char*)(((uintptr_t)p + kPiece - 1) / kPiece * kPiece);
    if (base > (char*)p) munmap(p, base - (char*)p);
    const size_t tail = ((char*)p + len + kPiece) - (base + len);
    if (tail) munmap(base + len, tail);
    madvise(base, len, MADV_HUGEPAGE);

    const auto nodes = numa_nodes_with_cpus();
    bool placed = false;
    if (policy == 1 && nodes.size() >= 2) {
#ifdef SYS_mbind
        unsigned long mask = 0;
        for (const auto& n : nodes) mask |= 1UL << n.first;
        if (!getenv("PSB200_NO_MBIND") &&
            syscall(SYS_mbind, base, len, 3UL /* MPOL_INTERLEAVE */, &mask, 65UL, 0UL) == 0) {
            touch_pieces(base, len, nodes, false);
            placed = true;
        }
#endif
        if (!placed) { touch_pieces(base, len, nodes, true); placed = true; }
    }
    if (!placed) memset(base, 0, len);                       // first touch by the caller: local pages

    int registered = 0;
    if (device_count() > 0) {
        const cudaError_t e = cudaHostRegister(base, len, cudaHostRegisterPortable);
        if (e != cudaSuccess) {
            cudaGetLastError();
            munmap(base, len);
            fail(e == cudaErrorMemoryAllocation ? ERR_OOM : ERR_CUDA, "host_alloc: cudaHostRegister: %s", cudaGetErrorString(e));
            return nullptr;
        }
        registered = 1;
    }
    std::lock_guard<std::mutex> lk(g_host_mutex);
    g_host_blocks[base] = HostBlock{len, registered};
    return base;
}

int psb200_host_free(void* p)
{
    if (!p) return OK;
    HostBlock b;
    {
        std::lock_guard<std::mutex> lk(g_host_mutex);
        auto it = g_host_blocks.find(p);
        if (it == g_host_blocks.end()) return fail(ERR_ARG, "host_free: not a psb200_host_alloc pointer");
        b = it->second;
        g_host_blocks.erase(it);
    }
    if (b.registered) cudaHostUnregister(p);
    munmap(p, b.bytes);
    return OK;
}

/* Where the pages of a psb200_host_alloc block are: counts[k] = sampled 2 MB pieces found on NUMA node k (k < maxnodes).
 * Returns the number of pieces sampled, 0 if the kernel does not answer (get_mempolicy filtered), -1 on a bad pointer. */
int psb200_host_placement(const void* p, int* counts, int maxnodes)
{
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_host_mutex);
        auto it = g_host_blocks.find((void*)p);
        if (it == g_host_blocks.end() || !counts || maxnodes < 1) { fail(ERR_ARG, "host_placement: bad arguments"); return -1; }
        bytes = it->second.bytes;
    }
    for (int k = 0; k < maxnodes; ++k) counts[k] = 0;
    // odd stride over the pieces and a page offset that varies with the piece, so that neither a piece-wise nor a
    // page-wise alternation can alias with the sampling
    const size_t npieces = bytes / kPiece, step = std::max<size_t>(1, npieces / 256) | 1;
    int seen = 0;
    for (size_t k = 0; k < npieces; k += step) {
        const int node = page_node((const char*)p + k * kPiece + (k % 512) * 4096);
        if (node < 0) return 0;
        if (node < maxnodes) counts[node]++;
        ++seen;
    }
    return seen;
}

/* NUMA nodes (with CPUs this process may use) that psb200_host_alloc(policy 1) spreads a block over. */
int psb200_host_numa_nodes(void) { return (int)numa_nodes_with_cpus().size(); }

}  // extern "C"
