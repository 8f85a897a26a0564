/* ref_runtime.cpp -- work-group scheduler of the GLSL-on-CPU harness (see glsl_emu.h).  TEST INFRASTRUCTURE ONLY. */
#include "glsl_emu.h"

namespace glsl
{
uvec3 gl_GlobalInvocationID, gl_LocalInvocationID, gl_WorkGroupID, gl_NumWorkGroups;

namespace
{
struct Fiber
{
  ucontext_t ctx;
  std::vector<char> stack;
  bool done = false;
  uvec3 local, global;
};
ucontext_t g_sched;
Fiber *g_cur = nullptr;
shader_main_fn g_fn = nullptr;

void fiber_entry()
{
  g_fn();
  g_cur->done = true;
  swapcontext(&g_cur->ctx, &g_sched);
}
} // namespace

void barrier()
{
  if (g_cur) /* yield to the scheduler; it resumes every invocation of the group in order */
    swapcontext(&g_cur->ctx, &g_sched);
}

void dispatch(shader_main_fn fn, uint gx, uint gy, uint gz, uint lx, uint ly, uint lz, bool with_barriers)
{
  gl_NumWorkGroups = uvec3(gx, gy, gz);
  const uint n_local = lx * ly * lz;
  std::vector<Fiber> fibers;
  if (with_barriers)
  {
    fibers.resize(n_local);
    for (auto &f : fibers)
      f.stack.resize(256 * 1024);
  }
  for (uint wz = 0; wz < gz; wz++)
    for (uint wy = 0; wy < gy; wy++)
      for (uint wx = 0; wx < gx; wx++)
      {
        gl_WorkGroupID = uvec3(wx, wy, wz);
        if (!with_barriers)
        {
          for (uint z = 0; z < lz; z++)
            for (uint y = 0; y < ly; y++)
              for (uint x = 0; x < lx; x++)
              {
                gl_LocalInvocationID = uvec3(x, y, z);
                gl_GlobalInvocationID = uvec3(wx * lx + x, wy * ly + y, wz * lz + z);
                fn();
              }
          continue;
        }
        g_fn = fn;
        uint i = 0;
        for (uint z = 0; z < lz; z++)
          for (uint y = 0; y < ly; y++)
            for (uint x = 0; x < lx; x++, i++)
            {
              Fiber &f = fibers[i];
              f.done = false;
              f.local = uvec3(x, y, z);
              f.global = uvec3(wx * lx + x, wy * ly + y, wz * lz + z);
              getcontext(&f.ctx);
              f.ctx.uc_stack.ss_sp = f.stack.data();
              f.ctx.uc_stack.ss_size = f.stack.size();
              f.ctx.uc_link = &g_sched;
              makecontext(&f.ctx, fiber_entry, 0);
            }
        bool any = true;
        while (any)
        {
          any = false;
          for (uint k = 0; k < n_local; k++)
          {
            Fiber &f = fibers[k];
            if (f.done)
              continue;
            any = true;
            g_cur = &f;
            gl_LocalInvocationID = f.local;
            gl_GlobalInvocationID = f.global;
            swapcontext(&g_sched, &f.ctx);
          }
        }
        g_cur = nullptr;
      }
}
} // namespace glsl
