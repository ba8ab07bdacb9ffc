// Run-time compilation of the code-specialised LMS_DEC kernel (lms_spec.cuh) for an arbitrary base matrix.
//
// The kernel text (frame_io.h + channel.cuh + lms_spec.cuh, embedded at build time as build/spec_src.inc) is
// prefixed to a generated `Code` struct holding the matrix as compile-time tables, compiled with NVRTC for the
// device's architecture (sm_100a on B200) and loaded through the CUDA runtime's library API.  NVRTC is opened
// with dlopen, so the engine has no link-time dependency on it: when it is missing, or the compilation fails,
// the caller is told and stays on the table-driven kernel (lms_fast.cu).  Compiled instances are cached per
// process, keyed by the generated text, so an SNR sweep compiles once.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "kernels.h"

namespace ldpcb200 {

namespace {

const char* const KERNEL_TEXT =
#include "spec_src.inc"
    ;

typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    void* so = nullptr;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    int (*DestroyProgram)(nvrtcProgram*) = nullptr;
    bool ok = false;
};

Nvrtc& nvrtc()
{
    static Nvrtc n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    const char* names[] = { "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so" };
    for (const char* nm : names)
        if ((n.so = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
    if (!n.so) return n;
#define SYM(f) *(void**)(&n.f) = dlsym(n.so, "nvrtc" #f)
    SYM(CreateProgram); SYM(CompileProgram); SYM(GetCUBINSize); SYM(GetCUBIN); SYM(GetProgramLogSize); SYM(GetProgramLog); SYM(DestroyProgram);
#undef SYM
    n.ok = n.CreateProgram && n.CompileProgram && n.GetCUBINSize && n.GetCUBIN && n.GetProgramLogSize && n.GetProgramLog && n.DestroyProgram;
    return n;
}

struct Compiled {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
};

std::mutex g_mu;
std::map<std::string, Compiled> g_cache;        // key: device ordinal + generated text

template <class V>
void put_array(std::ostringstream& o, const char* decl, const V& v, int n)
{
    o << decl << "[" << n << "] = { ";
    for (int i = 0; i < n; i++) o << (i ? ", " : "") << v[i];
    o << " };\n";
}

} // namespace

// The generated part of the translation unit: same text as tools/gen_lms_spec.py writes for the ahead-of-time instances.
std::string lms_spec_generate(const QcHost& g, int zp, int minb, int variant, int kind)
{
    std::ostringstream o;
    o << "namespace ldpcb200 { namespace gen_jit {\n";
    put_array(o, "__constant__ int RT_RP", g.rp, g.b + 1);
    put_array(o, "__constant__ int RT_COL", g.col, g.E);
    put_array(o, "__constant__ int RT_SH", g.sh, g.E);
    // tables of the tensor-memory kernel (lms_tmem.cuh; same as tools/gen_lms_spec.py tmem_tables): every block
    // column is kept in the rotation of its last writer
    std::vector<int> rot(g.c, 0), ri(g.c, 0), delta(g.E, 0), synsh(g.E, 0), lastw(g.E, 0);
    {
        std::vector<int> last(g.c, 0), cur, seen(g.c, 0);
        for (int e = 0; e < g.E; e++) last[g.col[e]] = g.sh[e];
        rot = last;
        cur = last;
        for (int e = 0; e < g.E; e++) {
            delta[e] = ((g.sh[e] - cur[g.col[e]]) % g.Z + g.Z) % g.Z;
            cur[g.col[e]] = g.sh[e];
            synsh[e] = ((g.sh[e] - rot[g.col[e]]) % g.Z + g.Z) % g.Z;
        }
        for (int k = 0; k < g.c; k++) ri[k] = (g.Z - rot[k]) % g.Z;
        for (int e = g.E - 1; e >= 0; e--)
            if (!seen[g.col[e]]) { lastw[e] = 1; seen[g.col[e]] = 1; }
    }
    int tcols = 32;
    while (tcols < g.E * ((zp / 32 + 3) / 4)) tcols *= 2;
    put_array(o, "__constant__ int RT_ROT", rot, g.c);
    put_array(o, "__constant__ int RT_RI", ri, g.c);
    put_array(o, "__constant__ int RT_SYNSH", synsh, g.E);
    o << "struct Code {\n";
    o << "    static constexpr int B = " << g.b << ", C = " << g.c << ", Z = " << g.Z << ", E = " << g.E << ", ZP = " << zp << ", MINB = " << minb
      << ", MAXDEG = " << g.maxdeg << ";\n";
    // variant 0: doubled columns, all check state in registers; 1: single copy, sign/position words in shared memory;
    // 2: columns in the rotation of their last writer, c2v messages in tensor memory (lms_tmem.cuh); 3: the same, two frames per CTA (lms_tmem2.cuh);
    // 4 / 5: IMS_DEC only, tensor memory, frames as fp16 pairs, one / two groups of zp threads per CTA (ims_h2.cuh)
    o << "    static constexpr bool DOUBLED = " << (variant != 1 ? "true" : "false") << ", PS_SMEM = " << (variant != 1 ? "false" : "true") << ";\n";
    o << "    static constexpr int TCOLS = " << tcols << ";\n";
    put_array(o, "    static constexpr int DELTA", delta, g.E);
    put_array(o, "    static constexpr int ROT", rot, g.c);
    put_array(o, "    static constexpr int RI", ri, g.c);
    put_array(o, "    static constexpr int SYNSH", synsh, g.E);
    put_array(o, "    static constexpr bool LAST", lastw, g.E);
    {   // column untouched by the previous block row: its value can be loaded one block row ahead (lms_tmem.cuh prefetch)
        std::vector<int> early(g.E, 0);
        for (int j = 1; j < g.b; j++)
            for (int e = g.rp[j]; e < g.rp[j + 1]; e++) {
                bool hit = false;
                for (int e2 = g.rp[j - 1]; e2 < g.rp[j]; e2++) hit |= g.col[e2] == g.col[e];
                early[e] = !hit;
            }
        put_array(o, "    static constexpr bool EARLY", early, g.E);
    }
    o << "    static __device__ __forceinline__ const int* rt_rot() { return RT_ROT; }\n";
    o << "    static __device__ __forceinline__ const int* rt_ri() { return RT_RI; }\n";
    o << "    static __device__ __forceinline__ const int* rt_synsh() { return RT_SYNSH; }\n";
    put_array(o, "    static constexpr int RP", g.rp, g.b + 1);
    put_array(o, "    static constexpr int COL", g.col, g.E);
    put_array(o, "    static constexpr int SH", g.sh, g.E);
    {   // first edge of its block column (ascending block rows): pass A of the flooding kernels stores instead of adding
        std::vector<int> first(g.E, 0), seen(g.c, 0);
        for (int e = 0; e < g.E; e++) { first[e] = !seen[g.col[e]]; seen[g.col[e]] = 1; }
        put_array(o, "    static constexpr bool FIRST", first, g.E);
    }
    o << "    static __device__ __forceinline__ const int* rt_rp() { return RT_RP; }\n";
    o << "    static __device__ __forceinline__ const int* rt_col() { return RT_COL; }\n";
    o << "    static __device__ __forceinline__ const int* rt_sh() { return RT_SH; }\n";
    o << "};\n} }\n";
    // kind 0: LMS_DEC (layered), 1: MS_DEC fp32 (flooding), 2: IMS_DEC (flooding, fixed point)
    if (kind == 0) {
        o << "extern \"C\" __global__ void __launch_bounds__(" << zp << ", " << minb << ") spec_jit(const __grid_constant__ ldpcb200::FrameIO io)\n";
        o << "{ ldpcb200::" << (variant == 3 ? "LmsTmem2" : variant == 2 ? "LmsTmem" : "LmsSpec") << "<ldpcb200::gen_jit::Code>::kernel(io); }\n";
    } else {
        const int groups = variant == 5 ? 2 : 1;          // ims_h2.cuh: groups of zp threads per CTA
        o << "extern \"C\" __global__ void __launch_bounds__(" << (variant >= 4 ? groups * zp : zp) << ", " << minb << ") spec_jit(const __grid_constant__ ldpcb200::FrameIO io, const ldpcb200::MsSpecParams sp)\n";
        if (variant >= 4) o << "{ ldpcb200::ImsH2<ldpcb200::gen_jit::Code, " << groups << ">::kernel(io, sp); }\n";      // IMS_DEC, frames as fp16 pairs (ims_h2.cuh)
        else o << "{ ldpcb200::" << (variant == 2 ? "MsTmem" : "MsSpec") << "<ldpcb200::gen_jit::Code, " << (kind == 2 ? "true" : "false") << ">::kernel(io, sp); }\n";
    }
    return o.str();
}

// generated text -> cubin for sm_<major><minor>[a]; needs no device
bool lms_spec_compile(const std::string& gen, int major, int minor, std::vector<char>& cubin, std::string& why)
{
    Nvrtc& n = nvrtc();
    if (!n.ok) { why = "NVRTC (libnvrtc.so.12) not found"; return false; }
    const std::string src = std::string(KERNEL_TEXT) + "\n" + gen;
    if (const char* dump = getenv("LDPCB200_JIT_DUMP")) {          // keep the translation unit for inspection (nvcc -Xptxas -v, cuobjdump)
        if (FILE* f = fopen(dump, "w")) { fputs(src.c_str(), f); fclose(f); }
    }
    nvrtcProgram prog = nullptr;
    if (n.CreateProgram(&prog, src.c_str(), "lms_spec_jit.cu", 0, nullptr, nullptr) != 0) { why = "nvrtcCreateProgram failed"; return false; }
    char arch[64];
    // B200 is compute capability 10.0 -> sm_100a (architecture-specific features allowed)
    snprintf(arch, sizeof arch, "--gpu-architecture=sm_%d%d%s", major, minor, major >= 9 ? "a" : "");
    std::vector<std::string> extra;                               // LDPCB200_JIT_DEFINES="-DA=1 -DB=0" (development: A/B of kernel variants)
    if (const char* d = getenv("LDPCB200_JIT_DEFINES")) {
        std::istringstream is(d);
        for (std::string tok; is >> tok;) extra.push_back(tok);
    }
    std::vector<const char*> opts = { arch, "--std=c++17", "-lineinfo", "-fmad=false" };
    for (const std::string& e : extra) opts.push_back(e.c_str());
    int rc = n.CompileProgram(prog, (int)opts.size(), opts.data());
    if (rc != 0) {
        size_t ls = 0;
        n.GetProgramLogSize(prog, &ls);
        std::string log(ls, '\0');
        if (ls) n.GetProgramLog(prog, &log[0]);
        why = "NVRTC compilation failed: " + log.substr(0, 2000);
        n.DestroyProgram(&prog);
        return false;
    }
    size_t cs = 0;
    n.GetCUBINSize(prog, &cs);
    cubin.resize(cs);
    n.GetCUBIN(prog, cubin.data());
    n.DestroyProgram(&prog);
    return true;
}

// Compile (or fetch from the cache) the specialised kernel for `g` on the current device.
// Returns nullptr and fills `why` when run-time compilation is not possible.
const void* lms_spec_jit(const QcHost& g, int zp, int minb, int variant, int kind, std::string& why)
{
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { why = "no device"; return nullptr; }
    const std::string gen = lms_spec_generate(g, zp, minb, variant, kind);
    const char* defs = getenv("LDPCB200_JIT_DEFINES");
    const std::string key = std::to_string(dev) + "\n" + (defs ? defs : "") + "\n" + gen;
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) return (const void*)it->second.kernel;
    std::vector<char> cubin;
    if (!lms_spec_compile(gen, prop.major, prop.minor, cubin, why)) return nullptr;
    Compiled c;
    cudaError_t e = cudaLibraryLoadData(&c.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&c.kernel, c.lib, "spec_jit");
    if (e != cudaSuccess) { why = std::string("loading the compiled kernel failed: ") + cudaGetErrorString(e); cudaGetLastError(); return nullptr; }
    g_cache[key] = c;
    return (const void*)c.kernel;
}

cudaError_t launch_lms_spec_jit(const void* kernel, int zp, size_t smem, const FrameIO& io, int grid, cudaStream_t s)
{
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    void* args[] = { (void*)&io };
    return cudaLaunchKernel(kernel, dim3(grid), dim3(zp), args, smem, s);
}

cudaError_t launch_ms_spec(const void* kernel, int zp, size_t smem, const FrameIO& io, const MsSpecParams& sp, int grid, cudaStream_t s)
{
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    void* args[] = { (void*)&io, (void*)&sp };
    return cudaLaunchKernel(kernel, dim3(grid), dim3(zp), args, smem, s);
}

} // namespace ldpcb200
