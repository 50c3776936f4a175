// Run-time specialised kernels: the B200 counterpart of the reference's `jax.jit` /
// `tf.function(jit_compile=True)` (core.py:1106-1107, :1069-1071).  The reference hands the traced operator to XLA,
// which generates one fused program per operator; here the host tracer (odil_b200/graph.py) emits CUDA C for the
// traced expression graph -- forward residual, loss partial sums, reverse-mode adjoint, Jacobian products -- and this
// file compiles it with NVRTC straight to an sm_100a cubin and launches it through the driver API.
//
// libnvrtc is opened with dlopen on first use (the rest of the library works without it) and the driver entry points
// come from cudaGetDriverEntryPoint, so libodil_b200.so links against libcudart only.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace odil {

namespace {

struct Nvrtc {
    void* handle = nullptr;
    nvrtcResult (*createProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    nvrtcResult (*destroyProgram)(nvrtcProgram*);
    nvrtcResult (*compileProgram)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*getProgramLogSize)(nvrtcProgram, size_t*);
    nvrtcResult (*getProgramLog)(nvrtcProgram, char*);
    nvrtcResult (*getCUBINSize)(nvrtcProgram, size_t*);
    nvrtcResult (*getCUBIN)(nvrtcProgram, char*);
    const char* (*getErrorString)(nvrtcResult);
};

Nvrtc* nvrtc() {
    static Nvrtc api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (api.handle) break;
    }
    if (!api.handle) return nullptr;
#define ODIL_SYM(field, sym)                                         \
    *(void**)(&api.field) = dlsym(api.handle, #sym);                 \
    if (!api.field) {                                                \
        api.handle = nullptr;                                        \
        return nullptr;                                              \
    }
    ODIL_SYM(createProgram, nvrtcCreateProgram)
    ODIL_SYM(destroyProgram, nvrtcDestroyProgram)
    ODIL_SYM(compileProgram, nvrtcCompileProgram)
    ODIL_SYM(getProgramLogSize, nvrtcGetProgramLogSize)
    ODIL_SYM(getProgramLog, nvrtcGetProgramLog)
    ODIL_SYM(getCUBINSize, nvrtcGetCUBINSize)
    ODIL_SYM(getCUBIN, nvrtcGetCUBIN)
    ODIL_SYM(getErrorString, nvrtcGetErrorString)
#undef ODIL_SYM
    return &api;
}

struct Driver {
    bool ok = false;
    CUresult (*moduleLoadData)(CUmodule*, const void*);
    CUresult (*moduleUnload)(CUmodule);
    CUresult (*moduleGetFunction)(CUfunction*, CUmodule, const char*);
    CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                             void**, void**);
    CUresult (*funcSetAttribute)(CUfunction, CUfunction_attribute, int);
    CUresult (*getErrorString)(CUresult, const char**);
};

Driver* driver() {
    static Driver d;
    static bool tried = false;
    if (tried) return d.ok ? &d : nullptr;
    tried = true;
    auto get = [](const char* name, void** out) {
        cudaDriverEntryPointQueryResult q;
        return cudaGetDriverEntryPoint(name, out, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess &&
               *out != nullptr;
    };
    d.ok = get("cuModuleLoadData", (void**)&d.moduleLoadData) && get("cuModuleUnload", (void**)&d.moduleUnload) &&
           get("cuModuleGetFunction", (void**)&d.moduleGetFunction) && get("cuLaunchKernel", (void**)&d.launchKernel) &&
           get("cuFuncSetAttribute", (void**)&d.funcSetAttribute) && get("cuGetErrorString", (void**)&d.getErrorString);
    return d.ok ? &d : nullptr;
}

std::string& jit_log_ref() {
    static thread_local std::string log;
    return log;
}

}  // namespace

struct JitModule {
    std::vector<char> cubin;
    CUmodule module = nullptr;
    bool loaded = false;
};

}  // namespace odil

using namespace odil;

extern "C" {

// Compiles CUDA C `source` to an sm_100a cubin (no device needed).  Extra NVRTC options may be passed.
int odil_b200_jit_compile(const char* source, const char* const* options, int noptions, void** module_out) {
    ODIL_REQUIRE(source && module_out, "null argument");
    Nvrtc* rt = nvrtc();
    ODIL_REQUIRE(rt != nullptr, "libnvrtc could not be loaded (needed for operators that are not affine stencils)");
    nvrtcProgram prog;
    nvrtcResult r = rt->createProgram(&prog, source, "odil_b200_generated.cu", 0, nullptr, nullptr);
    ODIL_REQUIRE(r == NVRTC_SUCCESS, "nvrtcCreateProgram: %s", rt->getErrorString(r));
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
    for (int i = 0; i < noptions; ++i) opts.push_back(options[i]);
    r = rt->compileProgram(prog, (int)opts.size(), opts.data());
    size_t logsz = 0;
    rt->getProgramLogSize(prog, &logsz);
    std::string& log = jit_log_ref();
    log.assign(logsz ? logsz : 1, '\0');
    if (logsz) rt->getProgramLog(prog, &log[0]);
    if (r != NVRTC_SUCCESS) {
        rt->destroyProgram(&prog);
        return fail("NVRTC compilation failed (%s); see odil_b200_jit_log()", rt->getErrorString(r));
    }
    size_t sz = 0;
    r = rt->getCUBINSize(prog, &sz);
    if (r != NVRTC_SUCCESS || sz == 0) {
        rt->destroyProgram(&prog);
        return fail("nvrtcGetCUBINSize: %s", rt->getErrorString(r));
    }
    JitModule* m = new JitModule;
    m->cubin.resize(sz);
    r = rt->getCUBIN(prog, m->cubin.data());
    rt->destroyProgram(&prog);
    if (r != NVRTC_SUCCESS) {
        delete m;
        return fail("nvrtcGetCUBIN: %s", rt->getErrorString(r));
    }
    *module_out = m;
    return 0;
}

const char* odil_b200_jit_log(void) { return jit_log_ref().c_str(); }

// The compiled image (for cuobjdump / caching).
int odil_b200_jit_cubin(void* module, const void** data, uint64_t* size) {
    ODIL_REQUIRE(module && data && size, "null argument");
    JitModule* m = (JitModule*)module;
    *data = m->cubin.data();
    *size = m->cubin.size();
    return 0;
}

// Kernel handle by name; loads the cubin into the current context on first use (needs a device).
int odil_b200_jit_kernel(void* module, const char* name, int max_dynamic_smem, void** kernel_out) {
    ODIL_REQUIRE(module && name && kernel_out, "null argument");
    JitModule* m = (JitModule*)module;
    Driver* d = driver();
    ODIL_REQUIRE(d != nullptr, "CUDA driver entry points are not available");
    if (!m->loaded) {
        ODIL_CUDA(cudaFree(0));  // make sure the primary context exists and is current
        CUresult r = d->moduleLoadData(&m->module, m->cubin.data());
        if (r != CUDA_SUCCESS) {
            const char* s = nullptr;
            d->getErrorString(r, &s);
            return fail("cuModuleLoadData: %s", s ? s : "?");
        }
        m->loaded = true;
    }
    CUfunction f;
    CUresult r = d->moduleGetFunction(&f, m->module, name);
    ODIL_REQUIRE(r == CUDA_SUCCESS, "kernel '%s' not found in the generated module", name);
    if (max_dynamic_smem > 48 * 1024) {
        r = d->funcSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, max_dynamic_smem);
        ODIL_REQUIRE(r == CUDA_SUCCESS, "cuFuncSetAttribute(max dynamic shared memory = %d) failed", max_dynamic_smem);
    }
    *kernel_out = (void*)f;
    return 0;
}

// Launches a generated kernel whose single parameter is a struct passed by value: `params` points at `nbytes` bytes.
int odil_b200_jit_launch(void* kernel, uint32_t grid, uint32_t block, uint32_t smem, const void* params, uint64_t nbytes,
                         void* stream) {
    ODIL_REQUIRE(kernel && params, "null argument");
    ODIL_REQUIRE(grid >= 1 && block >= 1 && block <= 1024, "bad launch geometry grid=%u block=%u", grid, block);
    ODIL_REQUIRE(nbytes <= 4096, "parameter block of %llu bytes exceeds the 4 KB kernel parameter space",
                 (unsigned long long)nbytes);
    Driver* d = driver();
    ODIL_REQUIRE(d != nullptr, "CUDA driver entry points are not available");
    size_t sz = (size_t)nbytes;
    void* extra[] = {CU_LAUNCH_PARAM_BUFFER_POINTER, (void*)params, CU_LAUNCH_PARAM_BUFFER_SIZE, &sz, CU_LAUNCH_PARAM_END};
    CUresult r = d->launchKernel((CUfunction)kernel, grid, 1, 1, block, 1, 1, smem, (CUstream)stream, nullptr, extra);
    launch_counter()++;
    if (r != CUDA_SUCCESS) {
        const char* s = nullptr;
        d->getErrorString(r, &s);
        return fail("cuLaunchKernel: %s", s ? s : "?");
    }
    return 0;
}

int odil_b200_jit_destroy(void* module) {
    if (!module) return 0;
    JitModule* m = (JitModule*)module;
    if (m->loaded) {
        Driver* d = driver();
        if (d) d->moduleUnload(m->module);
    }
    delete m;
    return 0;
}

}  // extern "C"
