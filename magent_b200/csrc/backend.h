// backend.h -- the thin seam between the host engine (engine.cc, plain C++) and the device.
//
// The product library links backend_cuda.cu (CUDA runtime + the sm_100a kernels).  tests/emu/ links
// backend_emu.cc instead, which runs the same phase functions single-threaded on the host so the
// parallel formulations can be debugged without a GPU.  There is exactly one backend per library;
// nothing selects between them at run time.
#pragma once
#include <stddef.h>
#include <string>
#include "dev_types.h"

namespace mg {
namespace be {

const char *name();
// Select `device` (or the current one when < 0).  Returns false and fills *err when no usable device.
bool init(int device, std::string *err);
int device_count();
int sm_count();

void *dmalloc(size_t bytes);
void dfree(void *p);
void dmemset(void *p, int byte, size_t bytes);
void h2d(void *dst, const void *src, size_t bytes);
void d2h(void *dst, const void *src, size_t bytes);
void d2d(void *dst, const void *src, size_t bytes);
void *host_alloc(size_t bytes);         // page-locked
void host_free(void *p);
bool is_device_ptr(const void *p);
void sync();

// dE: device copy of the EngineDev block; hE: the host copy it was uploaded from (for sizes)
void launch_step(const EngineDev *dE, const EngineDev &hE, const StepArgs &S, int max_agents_per_arena);
void launch_cull(const EngineDev *dE, const EngineDev &hE, unsigned curmask, int max_agents_per_arena);
void launch_offsets(const EngineDev *dE, const EngineDev &hE);
// per-call preparation of get_observation: normalised minimap into mm_val ([A][G][view cells]; nullptr when
// minimap_mode is off) plus whatever the backend wants to precompute for the render kernel
void launch_obs_prepare(const EngineDev *dE, const EngineDev &hE, unsigned curmask, int obs_group, float *mm_val);
// true while the backend-side products of the last launch_obs_prepare still belong to this engine block
bool obs_prepare_valid(const EngineDev *dE);
void launch_obs(const EngineDev *dE, const EngineDev &hE, const ObsArgs &O, const float *mm_val, int n_total);
void launch_info(const EngineDev *dE, const EngineDev &hE, unsigned curmask, int kind, int group,
                 void *buf, int n_total);
void launch_random_actions(const EngineDev *dE, const EngineDev &hE, unsigned curmask, int group,
                           unsigned long long seed, int n_total);

// instrumentation: kernel launch counter and optional CUDA-event timing of the obs-render kernel
long long launch_count();
void profile_enable(bool on);
void profile_read(double *obs_ms_total, long long *obs_launches);

}  // namespace be
}  // namespace mg
