// Per-model specialised pointwise kernels: load the cubin that mlx_mcmc_b200/jit.py compiled with NVRTC (the shared
// device headers with the model's term table baked in as literals) and launch its four entry points through the driver
// API.  The generic interpreter kernels of pointwise.cu / nuts_pointwise.cu stay the path for models without a module.
#include <cuda.h>

#include "pointwise.cuh"

namespace b2m {

struct JitModule {
  CUmodule mod = nullptr;
  CUfunction fn[4] = {nullptr, nullptr, nullptr, nullptr};   // logp_grad, hmc, mh, nuts
  int dmax = 0;
  size_t smem_set[4] = {0, 0, 0, 0};
};

namespace {

struct DriverApi {
  CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **,
                           void **) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  bool ok = false;
};

const DriverApi &driver() {
  static const DriverApi api = [] {   // resolved once (C++11 magic static)
    DriverApi a;
    auto get = [](const char *name) -> void * {
      void *p = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
      return p;
    };
    a.ModuleLoadData = reinterpret_cast<decltype(a.ModuleLoadData)>(get("cuModuleLoadData"));
    a.ModuleGetFunction = reinterpret_cast<decltype(a.ModuleGetFunction)>(get("cuModuleGetFunction"));
    a.ModuleUnload = reinterpret_cast<decltype(a.ModuleUnload)>(get("cuModuleUnload"));
    a.LaunchKernel = reinterpret_cast<decltype(a.LaunchKernel)>(get("cuLaunchKernel"));
    a.FuncSetAttribute = reinterpret_cast<decltype(a.FuncSetAttribute)>(get("cuFuncSetAttribute"));
    a.ok = a.ModuleLoadData && a.ModuleGetFunction && a.ModuleUnload && a.LaunchKernel && a.FuncSetAttribute;
    return a;
  }();
  return api;
}

}  // namespace

int jit_load(const void *image, int dmax, JitModule **out) {
  const DriverApi &d = driver();
  B2M_REQUIRE(d.ok, "jit: the CUDA driver entry points for module loading are not available");
  B2M_CHECK_CUDA(cudaFree(nullptr));   // make sure the runtime's primary context is current on this thread
  JitModule *m = new JitModule();
  m->dmax = dmax;
  CUresult r = d.ModuleLoadData(&m->mod, image);
  if (r != CUDA_SUCCESS) {
    delete m;
    set_error("jit: cuModuleLoadData failed with CUresult " + std::to_string((int)r) + " (cubin not built for this GPU?)");
    return 2;
  }
  static const char *names[4] = {"b2m_jit_logp_grad", "b2m_jit_hmc", "b2m_jit_mh", "b2m_jit_nuts"};
  for (int i = 0; i < 4; ++i) {
    r = d.ModuleGetFunction(&m->fn[i], m->mod, names[i]);
    if (r != CUDA_SUCCESS) {
      d.ModuleUnload(m->mod);
      delete m;
      set_error(std::string("jit: the module lacks the entry point ") + names[i]);
      return 2;
    }
  }
  *out = m;
  return 0;
}

void jit_unload(JitModule *m) {
  if (!m) return;
  if (m->mod && driver().ok) driver().ModuleUnload(m->mod);
  delete m;
}

int jit_dmax(const JitModule *m) { return m ? m->dmax : 0; }

int jit_launch(JitModule *m, int which, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void **params) {
  const DriverApi &d = driver();
  if (smem > 48 * 1024 && smem > m->smem_set[which]) {
    CUresult r = d.FuncSetAttribute(m->fn[which], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem);
    if (r != CUDA_SUCCESS) {
      set_error("jit: cuFuncSetAttribute(max dynamic shared memory) failed with CUresult " + std::to_string((int)r));
      return 2;
    }
    m->smem_set[which] = smem;
  }
  CUresult r = d.LaunchKernel(m->fn[which], grid.x, grid.y, grid.z, block.x, block.y, block.z, (unsigned)smem, st, params, nullptr);
  if (r != CUDA_SUCCESS) {
    set_error("jit: cuLaunchKernel failed with CUresult " + std::to_string((int)r));
    return 2;
  }
  ++g_launches;
  return 0;
}

}  // namespace b2m
