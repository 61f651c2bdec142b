/*
 * fav_lua_shim.c -- thin Lua C-API shim that lets the UNMODIFIED reference scripts
 * (fast_artistic_video.lua -> require 'stn' -> nn.BilinearSamplerBDHW) drive libfav_b200.so.
 *
 * It exports exactly what the reference's native libraries export:
 *   luaopen_libcustn   stnbdhw/init.cu:8-16, registering in the "nn" field of the torch.CudaTensor
 *                      metatable (BilinearSamplerBDHW.cu:195-200) the three names of :188-193
 *   luaopen_libstn     stnbdhw/init.c:11-24 (creates the global table `stn`; the BHWD CPU sampler it
 *                      registers there is dead code on this path: nn.BilinearSamplerBDHW has no CPU
 *                      backend in the reference either, utils.lua:145-147)
 * Call convention (BilinearSamplerBDHW.cu:111-152): stack 1 = module self, 2 = inputImages, 3 = grids,
 * 4 = output, each a torch.CudaTensor; sizes/strides read from the THCudaTensor, kernel enqueued on
 * THCState_getCurrentStream, no synchronisation, non-zero status -> THError (Lua error).
 *
 * Build (on a machine that has Torch7):   gcc -shared -fPIC -DFAV_WITH_LUA fav_lua_shim.c \
 *     -I$TORCH/include -I$TORCH/include/THC -L.. -lfav_b200 -lluaT -lTHC -o libcustn.so
 * This container has no Lua/Torch7 headers: without -DFAV_WITH_LUA the file compiles against the
 * minimal declarations below (syntax/ABI check only).
 */
#include <stdint.h>
#include <stdio.h>

#include "../../include/fav.h"

#ifdef FAV_WITH_LUA
#include "luaT.h"
#include "THC.h"
#else
/* ---- minimal stand-ins for lua.h / luaT.h / THC.h (declarations only) ---- */
typedef struct lua_State lua_State;
typedef int (*lua_CFunction)(lua_State *L);
typedef struct luaL_Reg { const char *name; lua_CFunction func; } luaL_Reg;
typedef struct THCState THCState;
typedef struct THCudaTensor { long *size; long *stride; int nDimension; } THCudaTensor;
typedef struct CUstream_st *cudaStream_t;
void *luaT_checkudata(lua_State *L, int ud, const char *tname);
int luaT_pushmetatable(lua_State *L, const char *tname);
void luaT_registeratname(lua_State *L, const struct luaL_Reg *methods, const char *name);
void lua_getglobal(lua_State *L, const char *name);
void lua_getfield(lua_State *L, int idx, const char *k);
void lua_call(lua_State *L, int nargs, int nresults);
void *lua_touserdata(lua_State *L, int idx);
void lua_settop(lua_State *L, int idx);
void lua_createtable(lua_State *L, int narr, int nrec);
void lua_pushvalue(lua_State *L, int idx);
void lua_setglobal(lua_State *L, const char *name);
#define lua_pop(L, n) lua_settop(L, -(n)-1)
#define lua_newtable(L) lua_createtable(L, 0, 0)
float *THCudaTensor_data(THCState *state, const THCudaTensor *t);
long THCudaTensor_size(THCState *state, const THCudaTensor *t, int dim);
long THCudaTensor_stride(THCState *state, const THCudaTensor *t, int dim);
cudaStream_t THCState_getCurrentStream(THCState *state);
void THError(const char *fmt, ...);
#define LUA_EXTERNC
#define DLL_EXPORT __attribute__((visibility("default")))
#endif

/* stnbdhw/utils.c:3-11 */
static THCState *fav_getCutorchState(lua_State *L) {
  lua_getglobal(L, "cutorch");
  lua_getfield(L, -1, "getState");
  lua_call(L, 0, 1);
  THCState *state = (THCState *)lua_touserdata(L, -1);
  lua_pop(L, 2);
  return state;
}

/* replaces cunn_BilinearSamplerBDHW_updateOutput (BilinearSamplerBDHW.cu:111-152) */
static int fav_BilinearSamplerBDHW_updateOutput(lua_State *L) {
  THCState *state = fav_getCutorchState(L);
  THCudaTensor *inputImages = (THCudaTensor *)luaT_checkudata(L, 2, "torch.CudaTensor");
  THCudaTensor *grids = (THCudaTensor *)luaT_checkudata(L, 3, "torch.CudaTensor");
  THCudaTensor *output = (THCudaTensor *)luaT_checkudata(L, 4, "torch.CudaTensor");
  int64_t isz[4], ist[4], gsz[4], gst[4], ost[4];
  for (int d = 0; d < 4; ++d) {
    isz[d] = THCudaTensor_size(state, inputImages, d);
    ist[d] = THCudaTensor_stride(state, inputImages, d);
    gsz[d] = THCudaTensor_size(state, grids, d);
    gst[d] = THCudaTensor_stride(state, grids, d);
    ost[d] = THCudaTensor_stride(state, output, d);
  }
  int st = fav_bilinear_sampler_bdhw_update_output(
      THCudaTensor_data(state, inputImages), isz, ist, THCudaTensor_data(state, grids), gsz, gst,
      THCudaTensor_data(state, output), ost, FAV_BORDER_PER_TAP, (void *)THCState_getCurrentStream(state));
  if (st != FAV_OK) { /* BilinearSamplerBDHW.cu:146-150 */
    printf("error in BilinearSampler.updateOutput: %s\n", fav_last_error());
    THError("aborting");
  }
  return 1;
}

/* BilinearSamplerBDHW.cu:171-176 */
static int fav_BilinearSamplerBDHW_updateGradInput(lua_State *L) {
  (void)L;
  fav_bilinear_sampler_bdhw_update_grad_input();
  printf("%s", fav_last_error());
  THError("aborting");
  return 0;
}

/* BilinearSamplerBDHW.cu:179-184 */
static int fav_BilinearSamplerBDHW_updateGradInputOnlyGrid(lua_State *L) {
  (void)L;
  fav_bilinear_sampler_bdhw_update_grad_input_only_grid();
  printf("%s", fav_last_error());
  THError("aborting");
  return 0;
}

/* BilinearSamplerBDHW.cu:188-193: same three names */
static const struct luaL_Reg fav_BilinearSamplerBDHW__[] = {
    {"BilinearSamplerBDHW_updateOutput", fav_BilinearSamplerBDHW_updateOutput},
    {"BilinearSamplerBDHW_updateGradInput", fav_BilinearSamplerBDHW_updateGradInput},
    {"BilinearSamplerBDHW_updateGradInputOnlyGrid", fav_BilinearSamplerBDHW_updateGradInputOnlyGrid},
    {NULL, NULL}};

LUA_EXTERNC DLL_EXPORT int luaopen_libcustn(lua_State *L);
LUA_EXTERNC DLL_EXPORT int luaopen_libstn(lua_State *L);

/* stnbdhw/init.cu:10-16 + BilinearSamplerBDHW.cu:195-200 */
int luaopen_libcustn(lua_State *L) {
  lua_newtable(L);
  luaT_pushmetatable(L, "torch.CudaTensor");
  luaT_registeratname(L, fav_BilinearSamplerBDHW__, "nn");
  lua_pop(L, 1);
  return 1;
}

/* stnbdhw/init.c:13-24: global table `stn`; the BHWD CPU sampler is not on the hot path and is not provided */
int luaopen_libstn(lua_State *L) {
  lua_newtable(L);
  lua_pushvalue(L, -1);
  lua_setglobal(L, "stn");
  return 1;
}
