/*
 * refcl — runs the REFERENCE'S OWN OpenCL kernels (kernels/octree.cl, mls.cl, marching.cl,
 * scale_bias.cl of bmerry/mlsgpu) on the box's OpenCL device.
 *
 * ORACLE — TEST INFRASTRUCTURE ONLY.  The kernel sources are NOT in this repository: the build
 * recipe (oracle/Makefile, target _ref/librefcl.so) embeds them with .incbin from where they lie
 * under /root/reference/kernels at build time, exactly like the reference's own build embeds them
 * as strings (utils/clc2cpp.py, wscript:392-396), and the output goes to oracle/_ref/ only.
 * Everything in THIS file is ours: the host side that the reference writes with Boost, clogs and
 * the OpenCL C++ bindings (none of which exist in the image) is restated in plain C on the OpenCL
 * C API, loaded with dlopen so that nothing links against OpenCL:
 *
 *   programs built as CLH::build does (src/clh.cpp:304-345): "#define K V" lines, #line, source
 *   octree:   SplatTreeCL::enqueueBuild (src/splat_tree_cl.cpp:269-335); clogs::Radixsort /
 *             clogs::Scan — third party, absent — replaced by a stable host radix sort on the low
 *             key bits and a host exclusive scan (their documented semantics)
 *   MLS:      MlsFunctor::enqueue (src/mls.cpp:101-135), swathe by swathe like Marching::generate
 *             (src/marching.cpp:745-824) because an OpenCL image cannot hold a whole 256^3 bin
 *   marching: generateCells / addSlices / shipOut for ONE swathe (src/marching.cpp:500-743)
 *
 * Used to pin the CPU oracle (and the CUDA path) against the real reference kernels, and to time
 * the reference's processCorners on the same B200 through NVIDIA's OpenCL.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CL_USE_DEPRECATED_OPENCL_1_1_APIS
#include <CL/cl.h>

#define API __attribute__((visibility("default")))

extern const char refcl_src_octree[], refcl_src_mls[], refcl_src_marching[], refcl_src_scale_bias[];

static char gErr[2048];
static void fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(gErr, sizeof(gErr), fmt, ap);
    va_end(ap);
}
API const char *refcl_error(void) { return gErr; }

/* ---- the OpenCL entry points we use, resolved with dlsym ---------------------------------- */
#define CLFUNCS(X) \
    X(clGetPlatformIDs) X(clGetDeviceIDs) X(clGetDeviceInfo) X(clGetPlatformInfo) X(clCreateContext) \
    X(clCreateCommandQueue) X(clCreateProgramWithSource) X(clBuildProgram) X(clGetProgramBuildInfo) \
    X(clCreateKernel) X(clSetKernelArg) X(clCreateBuffer) X(clCreateImage2D) X(clEnqueueWriteBuffer) \
    X(clEnqueueReadBuffer) X(clEnqueueWriteImage) X(clEnqueueReadImage) X(clEnqueueNDRangeKernel) \
    X(clFinish) X(clReleaseMemObject) X(clReleaseKernel) X(clReleaseProgram) X(clGetEventProfilingInfo) \
    X(clReleaseEvent) X(clWaitForEvents)
#define DECL(n) static __typeof__(&n) p_##n;
CLFUNCS(DECL)

static void *gLib;
static cl_context gCtx;
static cl_device_id gDev;
static cl_command_queue gQueue;
static char gDeviceName[256];
static size_t gMaxImageHeight, gMaxImageWidth;

#define CHECK(call)                                                        \
    do {                                                                   \
        cl_int e_ = (call);                                                \
        if (e_ != CL_SUCCESS) {                                            \
            fail("%s:%d: %s failed with %d", __FILE__, __LINE__, #call, e_); \
            return -1;                                                     \
        }                                                                  \
    } while (0)

API int refcl_open(void)
{
    if (gCtx) return 0;
    /* The box has NVIDIA's ICD library but no /etc/OpenCL/vendors entry: name it for the loader. */
    setenv("OCL_ICD_FILENAMES", "libnvidia-opencl.so.1", 0);
    const char *names[] = {"libOpenCL.so.1", "/usr/local/cuda/targets/x86_64-linux/lib/libOpenCL.so.1",
                           "/usr/local/cuda/lib64/libOpenCL.so.1", "libOpenCL.so"};
    for (unsigned i = 0; i < sizeof(names) / sizeof(names[0]) && !gLib; i++) gLib = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
    if (!gLib) { fail("no OpenCL ICD loader found: %s", dlerror()); return -1; }
#define LOAD(n)                                                             \
    p_##n = (__typeof__(p_##n)) dlsym(gLib, #n);                            \
    if (!p_##n) { fail("OpenCL symbol %s missing", #n); return -1; }
    CLFUNCS(LOAD)
    cl_platform_id plats[8];
    cl_uint np = 0;
    if (p_clGetPlatformIDs(8, plats, &np) != CL_SUCCESS || np == 0) { fail("no OpenCL platform (no ICD registered)"); return -1; }
    for (cl_uint i = 0; i < np && !gDev; i++)
    {
        cl_device_id d;
        cl_uint nd = 0;
        if (p_clGetDeviceIDs(plats[i], CL_DEVICE_TYPE_GPU, 1, &d, &nd) == CL_SUCCESS && nd > 0) gDev = d;
    }
    if (!gDev) { fail("no OpenCL GPU device"); return -1; }
    cl_int err;
    gCtx = p_clCreateContext(NULL, 1, &gDev, NULL, NULL, &err);
    if (!gCtx) { fail("clCreateContext failed with %d", err); return -1; }
    gQueue = p_clCreateCommandQueue(gCtx, gDev, CL_QUEUE_PROFILING_ENABLE, &err);   /* src/workers.cpp:215 */
    if (!gQueue) { fail("clCreateCommandQueue failed with %d", err); return -1; }
    p_clGetDeviceInfo(gDev, CL_DEVICE_NAME, sizeof(gDeviceName), gDeviceName, NULL);
    p_clGetDeviceInfo(gDev, CL_DEVICE_IMAGE2D_MAX_HEIGHT, sizeof(size_t), &gMaxImageHeight, NULL);
    p_clGetDeviceInfo(gDev, CL_DEVICE_IMAGE2D_MAX_WIDTH, sizeof(size_t), &gMaxImageWidth, NULL);
    return 0;
}

API const char *refcl_device_name(void) { return gDeviceName; }

/* CLH::build, src/clh.cpp:304-345 */
static cl_program buildProgram(const char *filename, const char *source, const char *defines)
{
    char header[1024];
    snprintf(header, sizeof(header), "%s#line 1 \"%s\"\n", defines, filename);
    const char *srcs[2] = {header, source};
    cl_int err;
    cl_program prog = p_clCreateProgramWithSource(gCtx, 2, srcs, NULL, &err);
    if (!prog) { fail("clCreateProgramWithSource(%s) failed with %d", filename, err); return NULL; }
    err = p_clBuildProgram(prog, 1, &gDev, "", NULL, NULL);
    if (err != CL_SUCCESS)
    {
        char log[1500] = "";
        p_clGetProgramBuildInfo(prog, gDev, CL_PROGRAM_BUILD_LOG, sizeof(log) - 1, log, NULL);
        fail("building %s failed with %d: %s", filename, err, log);
        p_clReleaseProgram(prog);
        return NULL;
    }
    return prog;
}

static double eventSeconds(cl_event ev)
{
    cl_ulong t0 = 0, t1 = 0;
    p_clWaitForEvents(1, &ev);
    p_clGetEventProfilingInfo(ev, CL_PROFILING_COMMAND_START, sizeof(t0), &t0, NULL);
    p_clGetEventProfilingInfo(ev, CL_PROFILING_COMMAND_END, sizeof(t1), &t1, NULL);
    p_clReleaseEvent(ev);
    return (double) (t1 - t0) * 1e-9;
}

/* CLH::enqueueNDRangeKernel: zero-sized ranges are skipped (src/clh.cpp) */
static int launch(cl_kernel k, cl_uint dims, const size_t *offset, const size_t *global, const size_t *local, double *seconds)
{
    for (cl_uint i = 0; i < dims; i++)
        if (global[i] == 0) return 0;
    cl_event ev;
    CHECK(p_clEnqueueNDRangeKernel(gQueue, k, dims, offset, global, local, 0, NULL, &ev));
    double s = eventSeconds(ev);
    if (seconds) *seconds += s;
    return 0;
}

#define ARG(k, i, v) CHECK(p_clSetKernelArg(k, i, sizeof(v), &(v)))

/* ======================================== octree ============================================ */

static cl_program gOctreeProg;
static cl_uint gOctreeLevels;

/* clogs::Radixsort(UINT key, INT value).enqueue(..., maxBits): stable, ascending, low bits */
static void stableSortPairs(uint32_t *keys, uint32_t *vals, size_t n, unsigned bits)
{
    uint32_t *k2 = malloc(n * 4), *v2 = malloc(n * 4);
    for (unsigned shift = 0; shift < bits; shift += 8)
    {
        unsigned w = bits - shift < 8 ? bits - shift : 8;
        uint32_t mask = (1u << w) - 1;
        size_t count[257] = {0};
        for (size_t i = 0; i < n; i++) count[((keys[i] >> shift) & mask) + 1]++;
        for (int d = 0; d < 256; d++) count[d + 1] += count[d];
        for (size_t i = 0; i < n; i++)
        {
            size_t dst = count[(keys[i] >> shift) & mask]++;
            k2[dst] = keys[i];
            v2[dst] = vals[i];
        }
        memcpy(keys, k2, n * 4);
        memcpy(vals, v2, n * 4);
    }
    free(k2);
    free(v2);
}

/* SplatTreeCL::enqueueBuild, src/splat_tree_cl.cpp:269-335.  splats (32 B each) are mutated like
 * the reference does (radius <- 1/r^2).  commands must hold 8n + 2*min(numStart, 8n) ints and is
 * NOT cleared (neither does the reference); keys/values (8n each, may be NULL) get the sorted
 * entries.  seconds accumulates the kernels' profiling time. */
API int refcl_octree_build(void *splats, uint32_t numSplatsTotal, uint32_t firstSplat, uint32_t numSplats,
                           const int32_t offset[3], uint32_t subsampling, uint32_t maxLevels, int32_t *start,
                           int32_t *commands, uint32_t *keysOut, uint32_t *valuesOut, double *seconds)
{
    if (refcl_open()) return -1;
    if (!gOctreeProg || gOctreeLevels != maxLevels)
    {
        char defs[64];
        snprintf(defs, sizeof(defs), "#define MAX_LEVELS %u\n", maxLevels);   /* src/splat_tree_cl.cpp:131-134 */
        if (gOctreeProg) p_clReleaseProgram(gOctreeProg);
        gOctreeProg = buildProgram("kernels/octree.cl", refcl_src_octree, defs);
        if (!gOctreeProg) return -1;
        gOctreeLevels = maxLevels;
    }
    cl_int err;
    const unsigned maxShift = maxLevels + subsampling - 1;
    const unsigned minShift = subsampling < maxShift ? subsampling : maxShift;
    size_t levelOffsets[40], pos = 0;
    for (unsigned i = minShift; i <= maxShift; i++)
    {
        levelOffsets[i] = pos;
        pos += (size_t) 1 << (3 * (maxShift - i));
    }
    const size_t numStart = pos, numEntries = (size_t) numSplats * 8;
    const size_t numCommands = numEntries + 2 * (numStart < numEntries ? numStart : numEntries);
    double secs = 0.0;
    if (numSplats == 0)
    {
        for (size_t i = 0; i < numStart; i++) start[i] = -1;
        return 0;
    }

    cl_kernel kWrite = p_clCreateKernel(gOctreeProg, "writeEntries", &err), kCount = p_clCreateKernel(gOctreeProg, "countCommands", &err),
              kIds = p_clCreateKernel(gOctreeProg, "writeSplatIds", &err), kFill = p_clCreateKernel(gOctreeProg, "fill", &err),
              kStart = p_clCreateKernel(gOctreeProg, "writeStart", &err), kTop = p_clCreateKernel(gOctreeProg, "writeStartTop", &err);
    if (!kWrite || !kCount || !kIds || !kFill || !kStart || !kTop) { fail("clCreateKernel(octree) failed with %d", err); return -1; }
    cl_mem dSplats = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, (size_t) numSplatsTotal * 32, splats, &err);
    cl_mem dKeys = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, numEntries * 4, NULL, &err);
    cl_mem dVals = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, numEntries * 4, NULL, &err);
    cl_mem dMap = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, numEntries * 4, NULL, &err);
    cl_mem dStart = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, numStart * 4, NULL, &err);
    cl_mem dJump = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, numStart * 4, NULL, &err);
    cl_mem dCmd = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, numCommands * 4, commands, &err);
    if (!dSplats || !dKeys || !dVals || !dMap || !dStart || !dJump || !dCmd) { fail("clCreateBuffer(octree) failed with %d", err); return -1; }

    /* enqueueWriteEntries, :143-173 */
    cl_int4 off4 = {{offset[0], offset[1], offset[2], 0}};
    cl_uint uMin = minShift, uMax = maxShift, uFirst = firstSplat;
    ARG(kWrite, 0, dKeys); ARG(kWrite, 1, dVals); ARG(kWrite, 2, dSplats); ARG(kWrite, 3, off4);
    CHECK(p_clSetKernelArg(kWrite, 4, sizeof(cl_uint) * (maxShift + 1), NULL));
    ARG(kWrite, 5, uMin); ARG(kWrite, 6, uMax); ARG(kWrite, 7, uFirst);
    size_t g = numSplats;
    if (launch(kWrite, 1, NULL, &g, NULL, &secs)) return -1;

    /* sort.enqueue(queue, entryKeys, entryValues, numEntries, 3 * (maxShift - minShift) + 1), :308 */
    uint32_t *hKeys = malloc(numEntries * 4), *hVals = malloc(numEntries * 4), *hMap = malloc(numEntries * 4);
    CHECK(p_clEnqueueReadBuffer(gQueue, dKeys, CL_TRUE, 0, numEntries * 4, hKeys, 0, NULL, NULL));
    CHECK(p_clEnqueueReadBuffer(gQueue, dVals, CL_TRUE, 0, numEntries * 4, hVals, 0, NULL, NULL));
    stableSortPairs(hKeys, hVals, numEntries, 3 * (maxShift - minShift) + 1);
    CHECK(p_clEnqueueWriteBuffer(gQueue, dKeys, CL_TRUE, 0, numEntries * 4, hKeys, 0, NULL, NULL));
    CHECK(p_clEnqueueWriteBuffer(gQueue, dVals, CL_TRUE, 0, numEntries * 4, hVals, 0, NULL, NULL));
    if (keysOut) memcpy(keysOut, hKeys, numEntries * 4);
    if (valuesOut) memcpy(valuesOut, hVals, numEntries * 4);

    /* enqueueCountCommands over numEntries - 1 items, :310 */
    ARG(kCount, 0, dMap); ARG(kCount, 1, dKeys);
    g = numEntries - 1;
    if (launch(kCount, 1, NULL, &g, NULL, &secs)) return -1;
    /* scan.enqueue(queue, commandMap, numEntries, &scanOffset = 1): exclusive, in place, :312-313 */
    CHECK(p_clEnqueueReadBuffer(gQueue, dMap, CL_TRUE, 0, numEntries * 4, hMap, 0, NULL, NULL));
    {
        uint32_t run = 1;
        for (size_t i = 0; i < numEntries; i++)
        {
            uint32_t v = hMap[i];
            hMap[i] = run;
            run += v;
        }
    }
    CHECK(p_clEnqueueWriteBuffer(gQueue, dMap, CL_TRUE, 0, numEntries * 4, hMap, 0, NULL, NULL));
    /* enqueueFill(jumpPos, 0, numStart, -1), :315 */
    cl_int minus1 = -1;
    ARG(kFill, 0, dJump); ARG(kFill, 1, minus1);
    g = numStart;
    if (launch(kFill, 1, NULL, &g, NULL, &secs)) return -1;
    /* enqueueWriteSplatIds, :317 */
    ARG(kIds, 0, dCmd); ARG(kIds, 1, dStart); ARG(kIds, 2, dJump); ARG(kIds, 3, dMap); ARG(kIds, 4, dKeys); ARG(kIds, 5, dVals);
    g = numEntries;
    if (launch(kIds, 1, NULL, &g, NULL, &secs)) return -1;
    /* writeStartTop, then writeStart coarse -> fine, :320-331 */
    for (int i = (int) maxShift; i >= (int) minShift; i--)
    {
        const int havePrev = i != (int) maxShift;
        cl_kernel k = havePrev ? kStart : kTop;
        cl_uint cur = (cl_uint) levelOffsets[i], prev = havePrev ? (cl_uint) levelOffsets[i + 1] : 0;
        ARG(k, 0, dStart); ARG(k, 1, dCmd); ARG(k, 2, dJump); ARG(k, 3, cur);
        if (havePrev) ARG(k, 4, prev);
        g = (size_t) 1 << (3 * (maxShift - i));
        if (launch(k, 1, NULL, &g, NULL, &secs)) return -1;
    }
    CHECK(p_clEnqueueReadBuffer(gQueue, dStart, CL_TRUE, 0, numStart * 4, start, 0, NULL, NULL));
    CHECK(p_clEnqueueReadBuffer(gQueue, dCmd, CL_TRUE, 0, numCommands * 4, commands, 0, NULL, NULL));
    CHECK(p_clEnqueueReadBuffer(gQueue, dSplats, CL_TRUE, 0, (size_t) numSplatsTotal * 32, splats, 0, NULL, NULL));
    free(hKeys); free(hVals); free(hMap);
    p_clReleaseMemObject(dSplats); p_clReleaseMemObject(dKeys); p_clReleaseMemObject(dVals); p_clReleaseMemObject(dMap);
    p_clReleaseMemObject(dStart); p_clReleaseMemObject(dJump); p_clReleaseMemObject(dCmd);
    p_clReleaseKernel(kWrite); p_clReleaseKernel(kCount); p_clReleaseKernel(kIds); p_clReleaseKernel(kFill);
    p_clReleaseKernel(kStart); p_clReleaseKernel(kTop);
    if (seconds) *seconds += secs;
    return 0;
}

/* ========================================= MLS ============================================== */

static cl_program gMlsProg[2];

static int mlsProgram(int shape)
{
    if (gMlsProg[shape]) return 0;
    char defs[256];
    /* src/mls.cpp:62-67 (std::map order: alphabetical) */
    snprintf(defs, sizeof(defs), "#define FIT_PLANE %d\n#define FIT_SPHERE %d\n#define WGS_X 8\n#define WGS_Y 8\n#define WGS_Z 8\n",
             shape == 1, shape == 0);
    gMlsProg[shape] = buildProgram("kernels/mls.cl", refcl_src_mls, defs);
    return gMlsProg[shape] ? 0 : -1;
}

/* MlsFunctor::enqueue (src/mls.cpp:101-135) for the corner slices zFirst..zLast of a region,
 * driven swathe by swathe as Marching::generate does (src/marching.cpp:786-797).  image is the
 * whole field, rows y + z*zStride + zBias with `pitch` floats per row, imageRows rows in all (the
 * kernel writes whole 8-slice work-groups, so slices up to the next multiple of 8 are returned
 * where the image has them); only the corners the kernel writes are changed.  maxSwathe = 0 picks the largest swathe
 * (a multiple of 8 slices) the device's image height allows. */
API int refcl_process_corners(float *image, uint32_t pitch, const void *splats, uint32_t numSplats, const int32_t *commands,
                              uint32_t numCommands, const int32_t *start, uint32_t numStart, uint32_t startShift,
                              const int32_t offset[3], uint32_t width, uint32_t height, uint32_t zStride, int32_t zBias,
                              uint32_t zFirst, uint32_t zLast, float boundaryFactor, int shape, uint32_t maxSwathe,
                              uint32_t imageRows, double *seconds, uint32_t *launches)
{
    if (refcl_open() || mlsProgram(shape)) return -1;
    cl_int err;
    const uint32_t w8 = (width + 7) / 8 * 8, h8 = (height + 7) / 8 * 8;
    if (zStride < h8 || pitch < w8 || zFirst % 8 != 0) { fail("bad swathe geometry"); return -1; }
    if (w8 > gMaxImageWidth) { fail("region wider than the device's images"); return -1; }
    uint32_t sw = (uint32_t) (gMaxImageHeight / zStride) / 8 * 8;
    if (maxSwathe && maxSwathe < sw) sw = maxSwathe / 8 * 8;
    if (sw == 0) { fail("one 8-slice swathe does not fit an image"); return -1; }
    const uint32_t slices = zLast - zFirst + 1;
    if (sw > (slices + 7) / 8 * 8) sw = (slices + 7) / 8 * 8;

    cl_kernel k = p_clCreateKernel(gMlsProg[shape], "processCorners", &err);
    if (!k) { fail("clCreateKernel(processCorners) failed with %d", err); return -1; }
    cl_image_format fmt = {CL_R, CL_FLOAT};   /* src/marching.cpp:380 */
    cl_mem dImg = p_clCreateImage2D(gCtx, CL_MEM_READ_WRITE, &fmt, w8, (size_t) zStride * sw, 0, NULL, &err);
    if (!dImg) { fail("clCreateImage2D failed with %d", err); return -1; }
    cl_mem dSplats = p_clCreateBuffer(gCtx, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, (size_t) (numSplats ? numSplats : 1) * 32, (void *) splats, &err);
    cl_mem dCmd = p_clCreateBuffer(gCtx, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, (size_t) (numCommands ? numCommands : 1) * 4, (void *) commands, &err);
    cl_mem dStart = p_clCreateBuffer(gCtx, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, (size_t) numStart * 4, (void *) start, &err);
    if (!dSplats || !dCmd || !dStart) { fail("clCreateBuffer(mls) failed with %d", err); return -1; }
    cl_int4 off4 = {{offset[0], offset[1], offset[2], 0}};
    cl_uint uShift = startShift, uStride = zStride;
    /* MlsFunctor::set (src/mls.cpp:76-88) + enqueue */
    ARG(k, 0, dImg); ARG(k, 1, dSplats); ARG(k, 2, dCmd); ARG(k, 3, dStart); ARG(k, 4, uShift); ARG(k, 5, off4);
    ARG(k, 6, uStride); ARG(k, 8, boundaryFactor);
    float *rows = malloc((size_t) w8 * zStride * sw * sizeof(float));
    double secs = 0.0;
    for (uint32_t z0 = zFirst; z0 <= zLast; z0 += sw)
    {
        const uint32_t z1 = z0 + sw - 1 < zLast ? z0 + sw - 1 : zLast;
        /* the swathe lives at the top of the image: zBias' = -z0 * zStride */
        cl_int bias = -(cl_int) (z0 * zStride);
        ARG(k, 7, bias);
        size_t nrows = (size_t) zStride * ((z1 - z0 + 8) / 8 * 8);
        const size_t firstRow = (size_t) z0 * zStride + zBias;
        if (firstRow + nrows > imageRows) nrows = imageRows - firstRow;
        const size_t origin[3] = {0, 0, 0}, region[3] = {w8, nrows, 1};
        /* preload what the caller has there, so that rows/columns the kernel does not write keep their value */
        for (size_t r = 0; r < region[1]; r++)
            memcpy(rows + r * w8, image + (firstRow + r) * pitch, w8 * sizeof(float));
        CHECK(p_clEnqueueWriteImage(gQueue, dImg, CL_TRUE, origin, region, 0, 0, rows, 0, NULL, NULL));
        const size_t goff[3] = {0, 0, z0}, global[3] = {512 * (size_t) (w8 / 8), h8 / 8, (z1 - z0 + 8) / 8}, local[3] = {512, 1, 1};
        if (launch(k, 3, goff, global, local, &secs)) return -1;
        if (launches) (*launches)++;
        CHECK(p_clEnqueueReadImage(gQueue, dImg, CL_TRUE, origin, region, 0, 0, rows, 0, NULL, NULL));
        for (size_t r = 0; r < region[1]; r++)
            memcpy(image + (firstRow + r) * pitch, rows + r * w8, w8 * sizeof(float));
    }
    free(rows);
    p_clReleaseMemObject(dImg); p_clReleaseMemObject(dSplats); p_clReleaseMemObject(dCmd); p_clReleaseMemObject(dStart);
    p_clReleaseKernel(k);
    if (seconds) *seconds += secs;
    return 0;
}

/* ======================================= marching =========================================== */

static cl_program gMarchProg, gScaleProg;

static int cmpKeyPos(const void *a, const void *b)
{
    const uint64_t *x = (const uint64_t *) a, *y = (const uint64_t *) b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return x[1] < y[1] ? -1 : x[1] > y[1];
}

/* clogs::Radixsort(ULONG key, float4 value), all 64 bits, stable: (key, original position) pairs
 * ordered by key, then position */
static void stableSortVertices(uint64_t *keys, float *verts4, size_t n)
{
    uint64_t *tmp = malloc(n * 16);
    for (size_t i = 0; i < n; i++) { tmp[2 * i] = keys[i]; tmp[2 * i + 1] = i; }
    qsort(tmp, n, 16, cmpKeyPos);
    float *v2 = malloc(n * 16);
    for (size_t i = 0; i < n; i++)
    {
        keys[i] = tmp[2 * i];
        memcpy(v2 + 4 * i, verts4 + 4 * tmp[2 * i + 1], 16);
    }
    memcpy(verts4, v2, n * 16);
    free(v2);
    free(tmp);
}

/* Marching::generate for a region that fits ONE swathe and one shipOut: generateCells (:500-551),
 * scanElements + generateElements (:721-731), shipOut (:553-625), then ScaleBiasFilter
 * (src/mesh_filter.cpp:92-113) when scale != 1 or bias != 0.  image: rows y + z*zStride (+ zBias),
 * pitch floats per row.  tables: count 256x2 u8, start 257x2 u16, data 8192 u8, key 2432x4 u32 (uint3
 * with 16-byte stride).  Outputs: vertices (packed xyz), keys (u64, valid from firstExternal on),
 * indices; sizes = {numWelded, numIndices, firstExternal, occupied cells, unwelded vertices}. */
API int refcl_marching_generate(const float *image, uint32_t pitch, uint32_t width, uint32_t height, uint32_t zStride,
                                int32_t zBias, uint32_t zFirst, uint32_t zLast, const uint32_t keyOffset[3],
                                const uint8_t *countTable, const uint16_t *startTable, const uint8_t *dataTable,
                                const uint32_t *keyTable4, float scale, const float bias[3], float *vertices,
                                uint64_t *keys, uint32_t *indices, uint32_t vertexCap, uint32_t indexCap, uint32_t sizes[5],
                                double *seconds)
{
    if (refcl_open()) return -1;
    if (!gMarchProg) gMarchProg = buildProgram("kernels/marching.cl", refcl_src_marching, "");
    if (!gScaleProg) gScaleProg = buildProgram("kernels/scale_bias.cl", refcl_src_scale_bias, "");
    if (!gMarchProg || !gScaleProg) return -1;
    cl_int err;
    double secs = 0.0;
    memset(sizes, 0, 5 * sizeof(uint32_t));
    const size_t imgRows = (size_t) zStride * (zLast + 1) + zBias;
    if (imgRows > gMaxImageHeight || width > gMaxImageWidth) { fail("region does not fit one image (%zu rows)", imgRows); return -1; }
    if (width < 2 || height < 2 || zLast <= zFirst) return 0;
    const size_t swatheCells = (size_t) (width - 1) * (height - 1) * (zLast - zFirst);

    cl_image_format fmt = {CL_R, CL_FLOAT};
    cl_mem dImg = p_clCreateImage2D(gCtx, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, &fmt, width, imgRows, (size_t) pitch * 4, (void *) image, &err);
    if (!dImg) { fail("clCreateImage2D(marching) failed with %d", err); return -1; }
    cl_mem dCells = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, swatheCells * 16, NULL, &err);
    cl_mem dVi = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, swatheCells * 8, NULL, &err);
    cl_uint zero = 0;
    cl_mem dN = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, 4, &zero, &err);
    uint32_t *hist = calloc((size_t) zLast + 1, 8);
    cl_mem dHist = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, ((size_t) zLast + 1) * 8, hist, &err);
    cl_mem dCount = p_clCreateBuffer(gCtx, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, 512, (void *) countTable, &err);
    cl_mem dStartT = p_clCreateBuffer(gCtx, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, 257 * 4, (void *) startTable, &err);
    cl_mem dData = p_clCreateBuffer(gCtx, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, 8192, (void *) dataTable, &err);
    cl_mem dKeyT = p_clCreateBuffer(gCtx, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, 2432 * 16, (void *) keyTable4, &err);
    if (!dCells || !dVi || !dN || !dHist || !dCount || !dStartT || !dData || !dKeyT) { fail("clCreateBuffer(marching) failed with %d", err); return -1; }

    /* ---- generateCells ---- */
    cl_kernel kOcc = p_clCreateKernel(gMarchProg, "genOccupied", &err);
    if (!kOcc) { fail("clCreateKernel(genOccupied) failed with %d", err); return -1; }
    cl_uint uStride = zStride;
    cl_int iBias = zBias;
    ARG(kOcc, 0, dCells); ARG(kOcc, 1, dVi); ARG(kOcc, 2, dN); ARG(kOcc, 3, dHist); ARG(kOcc, 4, dImg); ARG(kOcc, 5, uStride);
    ARG(kOcc, 6, iBias); ARG(kOcc, 7, dCount);
    {
        const size_t goff[3] = {0, 0, zFirst}, global[3] = {width - 1, height - 1, zLast - zFirst};
        if (launch(kOcc, 3, goff, global, NULL, &secs)) return -1;
    }
    cl_uint compacted = 0;
    CHECK(p_clEnqueueReadBuffer(gQueue, dN, CL_TRUE, 0, 4, &compacted, 0, NULL, NULL));
    sizes[3] = compacted;
    if (compacted == 0) goto done;
    {
        /* scanElements.enqueue(viCount, compacted, &offsets = (0, 0)): exclusive scan of uint2, :721 */
        uint32_t *vi = malloc((size_t) compacted * 8);
        CHECK(p_clEnqueueReadBuffer(gQueue, dVi, CL_TRUE, 0, (size_t) compacted * 8, vi, 0, NULL, NULL));
        uint32_t nv = 0, ni = 0;
        for (size_t i = 0; i < compacted; i++)
        {
            uint32_t v = vi[2 * i], x = vi[2 * i + 1];
            vi[2 * i] = nv;
            vi[2 * i + 1] = ni;
            nv += v;
            ni += x;
        }
        CHECK(p_clEnqueueWriteBuffer(gQueue, dVi, CL_TRUE, 0, (size_t) compacted * 8, vi, 0, NULL, NULL));
        free(vi);
        sizes[4] = nv;
        sizes[1] = ni;
        if (ni > indexCap) { fail("index capacity exceeded (%u)", ni); return -1; }

        cl_mem dUnw = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, (size_t) nv * 16, NULL, &err);
        cl_mem dUKeys = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, ((size_t) nv + 1) * 8, NULL, &err);
        cl_mem dIdx = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, (size_t) ni * 4, NULL, &err);
        cl_mem dUnique = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, ((size_t) nv + 1) * 4, NULL, &err);
        cl_mem dRemap = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, (size_t) nv * 4, NULL, &err);
        cl_mem dWelded = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, (size_t) nv * 16, NULL, &err);
        cl_mem dWKeys = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, (size_t) nv * 8, NULL, &err);
        cl_uint dead = 0xDEADBEEFu;
        cl_mem dFirstExt = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, 4, &dead, &err);
        if (!dUnw || !dUKeys || !dIdx || !dUnique || !dRemap || !dWelded || !dWKeys || !dFirstExt) { fail("clCreateBuffer(mesh) failed with %d", err); return -1; }

        /* ---- generateElements, :725-731; local size as src/marching.cpp:753-758 would pick for 48 KB ---- */
        cl_kernel kGen = p_clCreateKernel(gMarchProg, "generateElements", &err);
        if (!kGen) { fail("clCreateKernel(generateElements) failed with %d", err); return -1; }
        cl_uint4 gridOffset = {{keyOffset[0], keyOffset[1], keyOffset[2], 0}};
        cl_uint4 top = {{2 * (width - 1), 2 * (height - 1), 2 * zFirst, 0}};   /* src/marching.cpp:773-777 */
        const size_t localSize = 64;
        ARG(kGen, 0, dUnw); ARG(kGen, 1, dUKeys); ARG(kGen, 2, dIdx); ARG(kGen, 3, dVi); ARG(kGen, 4, dCells); ARG(kGen, 5, dImg);
        ARG(kGen, 6, dStartT); ARG(kGen, 7, dData); ARG(kGen, 8, dKeyT); ARG(kGen, 9, uStride); ARG(kGen, 10, iBias);
        ARG(kGen, 11, gridOffset); ARG(kGen, 12, top);
        CHECK(p_clSetKernelArg(kGen, 13, 19 * 16 * localSize, NULL));
        {
            /* enqueueNDRangeKernelSplit: whole work groups, then the remainder */
            size_t mainPart = compacted / localSize * localSize, rest = compacted - mainPart, zeroOff = 0;
            if (mainPart && launch(kGen, 1, &zeroOff, &mainPart, &localSize, &secs)) return -1;
            if (rest && launch(kGen, 1, &mainPart, &rest, &rest, &secs)) return -1;
        }
        /* ---- shipOut, :553-625 ---- */
        uint64_t *hKeys = malloc(((size_t) nv + 1) * 8);
        float *hVerts = malloc((size_t) nv * 16);
        CHECK(p_clEnqueueReadBuffer(gQueue, dUKeys, CL_TRUE, 0, (size_t) nv * 8, hKeys, 0, NULL, NULL));
        CHECK(p_clEnqueueReadBuffer(gQueue, dUnw, CL_TRUE, 0, (size_t) nv * 16, hVerts, 0, NULL, NULL));
        stableSortVertices(hKeys, hVerts, nv);
        hKeys[nv] = UINT64_MAX;   /* sentinel, :565-567 */
        CHECK(p_clEnqueueWriteBuffer(gQueue, dUKeys, CL_TRUE, 0, ((size_t) nv + 1) * 8, hKeys, 0, NULL, NULL));
        CHECK(p_clEnqueueWriteBuffer(gQueue, dUnw, CL_TRUE, 0, (size_t) nv * 16, hVerts, 0, NULL, NULL));
        free(hKeys);
        free(hVerts);
        cl_kernel kUniq = p_clCreateKernel(gMarchProg, "countUniqueVertices", &err), kCompact = p_clCreateKernel(gMarchProg, "compactVertices", &err),
                  kReindex = p_clCreateKernel(gMarchProg, "reindex", &err);
        if (!kUniq || !kCompact || !kReindex) { fail("clCreateKernel(weld) failed with %d", err); return -1; }
        ARG(kUniq, 0, dUnique); ARG(kUniq, 1, dUKeys);
        size_t g = nv;
        if (launch(kUniq, 1, NULL, &g, NULL, &secs)) return -1;
        /* scanUint over nv + 1 elements, no offset, :583 */
        uint32_t *hU = malloc(((size_t) nv + 1) * 4);
        CHECK(p_clEnqueueReadBuffer(gQueue, dUnique, CL_TRUE, 0, (size_t) nv * 4, hU, 0, NULL, NULL));
        uint32_t run = 0;
        for (size_t i = 0; i < nv; i++) { uint32_t v = hU[i]; hU[i] = run; run += v; }
        hU[nv] = run;
        CHECK(p_clEnqueueWriteBuffer(gQueue, dUnique, CL_TRUE, 0, ((size_t) nv + 1) * 4, hU, 0, NULL, NULL));
        free(hU);
        const uint32_t numWelded = run;
        sizes[0] = numWelded;
        if (numWelded > vertexCap) { fail("vertex capacity exceeded (%u)", numWelded); return -1; }
        cl_ulong minExternalKey = (cl_ulong) zLast << (2 * 21 + 1);
        cl_ulong keyOffsetL = ((cl_ulong) keyOffset[2] << (2 * 21 + 1)) | ((cl_ulong) keyOffset[1] << (21 + 1)) | ((cl_ulong) keyOffset[0] << 1);
        ARG(kCompact, 0, dWelded); ARG(kCompact, 1, dWKeys); ARG(kCompact, 2, dRemap); ARG(kCompact, 3, dFirstExt); ARG(kCompact, 4, dUnique);
        ARG(kCompact, 5, dUnw); ARG(kCompact, 6, dUKeys); ARG(kCompact, 7, minExternalKey); ARG(kCompact, 8, keyOffsetL);
        g = nv;
        if (launch(kCompact, 1, NULL, &g, NULL, &secs)) return -1;
        ARG(kReindex, 0, dIdx); ARG(kReindex, 1, dRemap);
        g = ni;
        if (launch(kReindex, 1, NULL, &g, NULL, &secs)) return -1;
        cl_uint firstExternal = 0;
        CHECK(p_clEnqueueReadBuffer(gQueue, dFirstExt, CL_TRUE, 0, 4, &firstExternal, 0, NULL, NULL));
        /* DeviceKeyMesh::assign: when no vertex is external the kernel never writes firstExternal */
        sizes[2] = firstExternal == 0xDEADBEEFu ? numWelded : firstExternal;
        /* ---- ScaleBiasFilter in place on the welded vertices ---- */
        if (scale != 1.0f || bias[0] != 0.0f || bias[1] != 0.0f || bias[2] != 0.0f)
        {
            cl_kernel kScale = p_clCreateKernel(gScaleProg, "scaleBiasVertices", &err);
            if (!kScale) { fail("clCreateKernel(scaleBiasVertices) failed with %d", err); return -1; }
            cl_float4 sb = {{bias[0], bias[1], bias[2], scale}};
            ARG(kScale, 0, dWelded); ARG(kScale, 1, sb);
            g = numWelded;
            if (launch(kScale, 1, NULL, &g, NULL, &secs)) return -1;
            p_clReleaseKernel(kScale);
        }
        CHECK(p_clEnqueueReadBuffer(gQueue, dWelded, CL_TRUE, 0, (size_t) numWelded * 12, vertices, 0, NULL, NULL));
        CHECK(p_clEnqueueReadBuffer(gQueue, dWKeys, CL_TRUE, 0, (size_t) numWelded * 8, keys, 0, NULL, NULL));
        CHECK(p_clEnqueueReadBuffer(gQueue, dIdx, CL_TRUE, 0, (size_t) ni * 4, indices, 0, NULL, NULL));
        p_clReleaseKernel(kGen); p_clReleaseKernel(kUniq); p_clReleaseKernel(kCompact); p_clReleaseKernel(kReindex);
        p_clReleaseMemObject(dUnw); p_clReleaseMemObject(dUKeys); p_clReleaseMemObject(dIdx); p_clReleaseMemObject(dUnique);
        p_clReleaseMemObject(dRemap); p_clReleaseMemObject(dWelded); p_clReleaseMemObject(dWKeys); p_clReleaseMemObject(dFirstExt);
    }
done:
    free(hist);
    p_clReleaseKernel(kOcc);
    p_clReleaseMemObject(dImg); p_clReleaseMemObject(dCells); p_clReleaseMemObject(dVi); p_clReleaseMemObject(dN); p_clReleaseMemObject(dHist);
    p_clReleaseMemObject(dCount); p_clReleaseMemObject(dStartT); p_clReleaseMemObject(dData); p_clReleaseMemObject(dKeyT);
    if (seconds) *seconds += secs;
    return 0;
}

/* The reference's UNIT_TESTS shim kernels of mls.cl (kernels/mls.cl:439-481) are compiled out
 * unless UNIT_TESTS is defined; build a second program with it for the scalar known-answer tests. */
static cl_program gMlsTestProg;
API int refcl_solve_quadratic(float a, float b, float c, float *out)
{
    if (refcl_open()) return -1;
    if (!gMlsTestProg)
        gMlsTestProg = buildProgram("kernels/mls.cl", refcl_src_mls,
                                    "#define FIT_PLANE 0\n#define FIT_SPHERE 1\n#define UNIT_TESTS 1\n#define WGS_X 8\n#define WGS_Y 8\n#define WGS_Z 8\n");
    if (!gMlsTestProg) return -1;
    cl_int err;
    cl_kernel k = p_clCreateKernel(gMlsTestProg, "testSolveQuadratic", &err);
    if (!k) { fail("clCreateKernel(testSolveQuadratic) failed with %d", err); return -1; }
    cl_mem d = p_clCreateBuffer(gCtx, CL_MEM_READ_WRITE, 4, NULL, &err);
    ARG(k, 0, d); ARG(k, 1, a); ARG(k, 2, b); ARG(k, 3, c);
    size_t one = 1;
    if (launch(k, 1, NULL, &one, NULL, NULL)) return -1;
    CHECK(p_clEnqueueReadBuffer(gQueue, d, CL_TRUE, 0, 4, out, 0, NULL, NULL));
    p_clReleaseMemObject(d);
    p_clReleaseKernel(k);
    return 0;
}
