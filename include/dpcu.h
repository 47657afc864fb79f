/* dpcu.h - C ABI of the B200-native culling backend (libdpcu.so)
 *
 * This is the drop-in boundary for nvpro-pipeline's data-parallel culling hot path.  Host code
 * (the C++ dp::culling::cuda::Manager in pipeline_b200/dp/culling/cuda, or any other FFI
 * client) calls ONLY these functions; nothing above this line knows about CUDA types and
 * nothing below it knows about dp:: types.  Every entry point names the reference interface
 * it replaces as path:line relative to the reference tree.
 *
 * Conventions
 *   - every function returns an int status: 0 (DPCU_OK) on success, otherwise a DPCU_ERR_* code
 *     or, for CUDA runtime failures, the positive cudaError_t value.  dpcuGetLastError() returns
 *     a thread-local, human readable message for the last failure on the calling thread.  The
 *     C++ layer turns non-zero into std::runtime_error exactly like CUDA_VERIFY does
 *     (dp/cuda/Config.h:47-57).
 *   - matrices are 16 consecutive floats, row-major, row-vector convention
 *     (dp/math/Matmnt.h:1371-1379); a "float4 array" is n * 4 floats, 16-byte aligned.
 *   - visibility bitsets use the reference layout: object i -> u32 word i/32, bit i%32, unused
 *     tail bits 0 (dp/util/BitArray.h:215-219,298-308).
 *   - memspace says where a caller pointer lives: DPCU_MEM_HOST (pageable or pinned host
 *     memory; the call copies before returning, so borrowed pointers may be released after
 *     the call exactly as with GroupBitSet::setMatrices + cull, dp/culling/src/GroupBitSet.cpp:121-135)
 *     or DPCU_MEM_DEVICE (memory of the context's device; copied device-to-device).
 *   - there is NO CPU fallback: every entry point fails with DPCU_ERR_NO_DEVICE when no CUDA
 *     device is usable.
 *   - a context is single-threaded like the reference Manager (SURVEY.md 8b "Threading"); use
 *     one context per thread or serialise externally.
 */
#ifndef DPCU_H
#define DPCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPCU_OK                 0
#define DPCU_ERR_INVALID_VALUE  (-1)   /* bad argument (null handle, index out of range, size mismatch) */
#define DPCU_ERR_NO_DEVICE      (-2)   /* no usable CUDA device: there is no CPU fallback               */
#define DPCU_ERR_OUT_OF_MEMORY  (-3)
#define DPCU_ERR_NOT_READY      (-4)
#define DPCU_ERR_UNSUPPORTED    (-5)

#define DPCU_MEM_HOST    0
#define DPCU_MEM_DEVICE  1

#define DPCU_MAX_VIEWS   8             /* views handled by one pass over the objects */

const char *dpcuGetLastError(void);
/* library version, e.g. 0x00010000 */
int dpcuGetVersion(void);

/* ===================================================================== dp/cuda layer
 * Thin C equivalents of the host RAII wrappers in dp/cuda (Buffer, BufferHost, Stream, Event,
 * Device).  GL interop (GraphicsResource) and pitched 3-D buffers (Buffer3D) are not needed by
 * the culling path and are not provided. */
typedef struct dpcuBuffer     dpcuBuffer;      /* dp::cuda::Buffer      dp/cuda/Buffer.h:39-76         */
typedef struct dpcuHostBuffer dpcuHostBuffer;  /* dp::cuda::BufferHost  dp/cuda/BufferHost.h:38-60     */
typedef struct dpcuStream     dpcuStream;      /* dp::cuda::Stream      dp/cuda/Stream.h:38-62         */
typedef struct dpcuEvent      dpcuEvent;       /* dp::cuda::Event       dp/cuda/Event.h:38-64          */

/* dp::cuda::Device (dp/cuda/Device.h:39-60, src/Device.cpp).  Unlike the reference's Device
 * destructor nothing here ever calls cudaDeviceReset. */
int dpcuDeviceCount(int *count);
int dpcuDeviceSelect(int device);
int dpcuDeviceCurrent(int *device);
int dpcuDeviceSynchronize(void);
/* name: caller buffer of nameBytes; any out pointer may be NULL */
int dpcuDeviceInfo(int device, char *name, size_t nameBytes, int *smCount, size_t *globalMemBytes,
                   int *ccMajor, int *ccMinor);

/* dp::cuda::Buffer::create / ~Buffer / setData / getData / fill (dp/cuda/src/Buffer.cpp) */
int dpcuBufferCreate(dpcuBuffer **out, size_t bytes);
int dpcuBufferDestroy(dpcuBuffer *buffer);
int dpcuBufferSize(const dpcuBuffer *buffer, size_t *bytes);
int dpcuBufferDevicePointer(const dpcuBuffer *buffer, void **devicePointer);
/* stream may be NULL: synchronous copy, like the reference's setData(void const*, size_t) */
int dpcuBufferUpload(dpcuBuffer *buffer, size_t offset, const void *host, size_t bytes, dpcuStream *stream);
int dpcuBufferDownload(const dpcuBuffer *buffer, size_t offset, void *host, size_t bytes, dpcuStream *stream);
int dpcuBufferFill(dpcuBuffer *buffer, int byteValue, size_t bytes, size_t offset);
/* the same, ordered on `stream` (cudaMemsetAsync) */
int dpcuBufferFillAsync(dpcuBuffer *buffer, int byteValue, size_t bytes, size_t offset, dpcuStream *stream);

/* dp::cuda::BufferHost::create (pinned, optionally mapped/portable; dp/cuda/src/BufferHost.cpp) */
#define DPCU_HOST_DEFAULT   0u
#define DPCU_HOST_PORTABLE  1u
#define DPCU_HOST_MAPPED    2u
#define DPCU_HOST_WRITECOMBINED 4u
int dpcuHostBufferCreate(dpcuHostBuffer **out, size_t bytes, unsigned flags);
int dpcuHostBufferDestroy(dpcuHostBuffer *buffer);
int dpcuHostBufferPointer(const dpcuHostBuffer *buffer, void **hostPointer);
int dpcuHostBufferSize(const dpcuHostBuffer *buffer, size_t *bytes);

/* dp::cuda::Stream (dp/cuda/src/Stream.cpp): blocking flag, priority, synchronize, wait(event) */
int dpcuStreamCreate(dpcuStream **out, int blocking, int priority);
int dpcuStreamDestroy(dpcuStream *stream);
int dpcuStreamSynchronize(dpcuStream *stream);
int dpcuStreamIsCompleted(dpcuStream *stream, int *completed);
int dpcuStreamWaitEvent(dpcuStream *stream, dpcuEvent *event);
/* raw cudaStream_t for interop (e.g. torch.cuda.ExternalStream) */
int dpcuStreamNative(dpcuStream *stream, void **cudaStream);

/* dp::cuda::Event (dp/cuda/src/Event.cpp) and getElapsedTime (dp/cuda/Event.h:62) */
#define DPCU_EVENT_DEFAULT         0u
#define DPCU_EVENT_BLOCKING_SYNC   1u
#define DPCU_EVENT_DISABLE_TIMING  2u
int dpcuEventCreate(dpcuEvent **out, unsigned flags);
int dpcuEventDestroy(dpcuEvent *event);
int dpcuEventRecord(dpcuEvent *event, dpcuStream *stream);
int dpcuEventSynchronize(dpcuEvent *event);
int dpcuEventIsCompleted(dpcuEvent *event, int *completed);
int dpcuEventElapsedMs(dpcuEvent *start, dpcuEvent *stop, float *milliseconds);

/* ===================================================================== culling layer
 * One dpcuCull is the device mirror of one culling group (dp::culling::GroupBitSet,
 * dp/culling/GroupBitSet.h:40-104): the object array as SoA float4 streams plus the world
 * matrices.  One dpcuCullResult is the device mirror of one ResultBitSet
 * (dp/culling/ResultBitSet.h:38-76): visibility bits of the previous cull plus the list of
 * objects whose visibility changed in the last cull. */
typedef struct dpcuCull       dpcuCull;
typedef struct dpcuCullResult dpcuCullResult;
typedef struct dpcuTree       dpcuTree;        /* transform layer, below */

/* replaces cpu::Manager::create + groupCreate (dp/culling/cpu/src/ManagerImpl.cpp:171-189) */
int dpcuCullCreate(dpcuCull **out, int device);
int dpcuCullDestroy(dpcuCull *ctx);

/* Replace the whole object array (the m_inputChanged upload, dp/culling/opengl/src/GroupImpl.cpp:75-113;
 * values as produced by ManagerBitSet::objectSetBoundingBox / objectSetTransformIndex,
 * dp/culling/src/ManagerBitSet.cpp:88-109):
 *   lower4[i]  = (box.lower.xyz, ignored)      extent4[i] = (box.upper - box.lower, ignored)
 *   transformIndex[i] = index into the matrix array; may be NULL when memspace is DEVICE and
 *   lower4[i].w already holds the index as raw u32 bits (the packed device layout).
 * Object i of this array is group index i.  n may be 0. */
int dpcuCullSetObjects(dpcuCull *ctx, const float *lower4, const float *extent4,
                       const uint32_t *transformIndex, size_t n, int memspace);
/* overwrite objects [first, first+count) of the current array (objectSetBoundingBox on live objects) */
int dpcuCullSetObjectRange(dpcuCull *ctx, size_t first, size_t count, const float *lower4,
                           const float *extent4, const uint32_t *transformIndex, int memspace);
int dpcuCullGetObjectCount(const dpcuCull *ctx, size_t *n);
/* Batched edits of live objects - what a frame of objectSetBoundingBox / objectSetTransformIndex /
 * groupAddObject / groupRemoveObject (GroupBitSet.cpp:76-119: the last object moves into the freed slot)
 * amounts to on the device: dpcuCullSetObjectCount grows or shrinks the object arrays KEEPING their
 * contents, dpcuCullUpdateObjects overwrites objects indices[k] with (lower4[k], extent4[k],
 * transformIndex[k]) for k < n (host arrays; indices >= the object count are skipped).  One staged
 * copy and one kernel per call; the caller does not wait for them. */
int dpcuCullSetObjectCount(dpcuCull *ctx, size_t n);
int dpcuCullUpdateObjects(dpcuCull *ctx, const uint32_t *indices, size_t n, const float *lower4, const float *extent4,
                          const uint32_t *transformIndex);

/* groupSetMatrices (dp/culling/src/GroupBitSet.cpp:121-135): copy `count` matrices laid out with
 * `strideBytes` between them (>= 64, multiple of 4) into the context.  The pointer is not retained. */
int dpcuCullSetMatrices(dpcuCull *ctx, const void *matrices, size_t count, size_t strideBytes, int memspace);
/* groupMatrixChanged (dp/culling/GroupBitSet.h:140-150) for a batch: re-read matrices[indices[k]]
 * (same base/stride addressing as above) for k < n; indices >= the current matrix count are
 * ignored like markMatrixDirty does.  memspace describes `matrices`; `indices` is host memory. */
int dpcuCullUpdateMatrices(dpcuCull *ctx, const uint32_t *indices, size_t n, const void *matrices,
                           size_t strideBytes, int memspace);
/* Zero-copy feed: cull straight out of a device matrix array owned by someone else (SURVEY.md
 * section 7 hard part 4).  64-byte stride, 32-byte aligned (the kernels read a matrix as two
 * 256-bit loads; anything cudaMalloc returns is).  The memory must stay valid until the context is
 * rebound, given its own copy, or destroyed.  ORDERING IS THE CALLER'S: whatever writes these
 * matrices must be ordered before every dpcuCullRun that reads them, and after the previous one,
 * by running on the same stream or through events (dpcuStreamWaitEvent).  For the world matrices
 * of a dpcuTree use dpcuCullBindTree, which does that ordering itself. */
int dpcuCullBindMatrices(dpcuCull *ctx, const void *deviceMatrices, size_t count);
/* Zero-copy feed out of a dpcuTree's world matrices (xbar: TransformTree -> culling,
 * dp/sg/xbar/culling/src/CullingImpl.cpp:126-144 without the host round trip).  Every dpcuCullRun
 * waits for the tree's last dpcuTreeCompute and every dpcuTreeCompute waits for the last cull that
 * read the matrices, on whatever streams they run (events, no host block); the binding follows the
 * tree when dpcuTreeSetTopology reallocates its arrays.  Unbind (dpcuCullBindMatrices /
 * dpcuCullSetMatrices / another tree) or destroy the context before destroying the tree. */
int dpcuCullBindTree(dpcuCull *ctx, dpcuTree *tree);
int dpcuCullGetMatrixCount(const dpcuCull *ctx, size_t *count);

/* groupCreateResult (dp/culling/cpu/src/ManagerImpl.cpp:191-194) */
int dpcuCullResultCreate(dpcuCull *ctx, dpcuCullResult **out);
int dpcuCullResultDestroy(dpcuCullResult *result);

/* Manager::cull (dp/culling/cpu/src/ManagerImpl.cpp:467-518 + ResultBitSet::updateChanged,
 * dp/culling/src/ResultBitSet.cpp:61-108) for nViews results at once: one pass over the objects,
 * view v tested against viewProjections[16*v .. 16*v+16) and recorded in results[v].
 * 1 <= nViews <= DPCU_MAX_VIEWS, results distinct and created from ctx.  On the first cull of a
 * result, and after the object count changed, bits of new objects start as 1 (visible) exactly
 * like ResultBitSet.cpp:65-79, so the first changed list is the set of invisible objects.
 * Asynchronous on `stream` (NULL = the context's own stream); the result getters synchronise. */
int dpcuCullRun(dpcuCull *ctx, dpcuCullResult *const *results, const float *viewProjections,
                int nViews, dpcuStream *stream);

/* visibility words of the last cull, nWords >= ceil(n/32) */
int dpcuCullResultGetBits(dpcuCullResult *result, uint32_t *hostWords, size_t nWords);
/* resultGetChanged (dp/culling/src/ManagerBitSet.cpp:141-144): group indices whose visibility
 * changed in the last cull, ASCENDING (BitArray::traverseBits order, dp/util/BitArray.h:127-136) */
int dpcuCullResultGetChangedCount(dpcuCullResult *result, size_t *count);
int dpcuCullResultGetChanged(dpcuCullResult *result, uint32_t *hostIndices, size_t capacity, size_t *count);
/* resultObjectIsVisible (dp/culling/ResultBitSet.h:69-76): true when index >= result size */
int dpcuCullResultIsVisible(dpcuCullResult *result, size_t groupIndex, int *visible);
/* ResultBitSet::onNotify (dp/culling/src/ResultBitSet.cpp:110-128): the group moved an object
 * from oldIndex to newIndex (GroupBitSet::removeObject, dp/culling/src/GroupBitSet.cpp:93-119) */
int dpcuCullResultMoveBit(dpcuCullResult *result, size_t oldIndex, size_t newIndex);
/* device-resident outputs for consumers that stay on the GPU and for the
 * multi-GPU gather.  *changedCount points at one u32. Valid until the next run / destroy. */
int dpcuCullResultDevicePointers(dpcuCullResult *result, const uint32_t **bits, size_t *nWords,
                                 const uint32_t **changedIndices, const uint32_t **changedCount);

/* GPU-side consumer of the result (replaces the host loop over changed objects in
 * DrawableManagerDefault::cull, dp/sg/renderer/rix/gl/src/DrawableManagerDefault.cpp:351-378, for renderers
 * that draw from the GPU): build, on the device, the ASCENDING list of group indices that are visible
 * after the last cull.  *count (one u32 in device memory) can be bound as the instance count of an
 * indirect draw, indices as the instance -> object table.  Valid until the next cull on this result;
 * moved bits (dpcuCullResultMoveBit) are reflected by building again. */
int dpcuCullResultBuildVisibleList(dpcuCullResult *result, dpcuStream *stream);
int dpcuCullResultVisibleDevicePointers(dpcuCullResult *result, const uint32_t **indices, const uint32_t **count);
int dpcuCullResultGetVisible(dpcuCullResult *result, uint32_t *hostIndices, size_t capacity, size_t *count);

/* Result mirror in pinned host memory.  The reference's Result lives in host memory
 * (ResultBitSet::m_results / m_changedObjects, dp/culling/ResultBitSet.h:57-61); with a mirror the
 * cull writes the visibility words and the changed list straight into the caller's pinned,
 * device-mapped buffers (dpcuHostBufferCreate) over PCIe while it runs - whole 128-byte bitset lines
 * and coalesced runs of the changed list from the line-granular cull kernel (large groups), or a copy /
 * the compaction kernel's stores queued behind the other kernel forms - so that after
 * dpcuCullResultSynchronize the result is readable on the host without any further call:
 *   hostBits[0 .. ceil(n/32))            visibility words (nWords = capacity, checked by dpcuCullRun)
 *   *hostChangedCount                    length of the changed list
 *   hostChanged[0 .. min(count, cap))    ascending group indices
 * hostBits and the hostChanged/hostChangedCount pair are optional independently; all NULL
 * removes the mirror.  dpcuCullResultIsVisible / MoveBit keep the mirror current.  The buffers
 * must stay valid until the mirror is changed or the result destroyed. */
int dpcuCullResultSetHostMirror(dpcuCullResult *result, uint32_t *hostBits, size_t nWords,
                                uint32_t *hostChanged, size_t changedCapacity, uint32_t *hostChangedCount);
/* block the calling thread until the last cull / bit move submitted for this result is complete */
int dpcuCullResultSynchronize(dpcuCullResult *result);
/* Overwrite visibility words of the stored result: words[k] -> word indices[k] (host arrays).  For hosts that
 * keep a mirror of the bits and apply a frame's bit moves (ResultBitSet::onNotify) there: the touched words
 * go back in one batch before the next cull instead of one dpcuCullResultMoveBit launch per removed object. */
int dpcuCullResultUpdateWords(dpcuCullResult *result, const uint32_t *indices, const uint32_t *words, size_t n);

/* ManagerBitSet::getBoundingBox / calculateBoundingBox, scalar branch
 * (dp/culling/src/ManagerBitSet.cpp:151-162,268-306): out6 = lower.xyz, upper.xyz */
int dpcuCullGetBoundingBox(dpcuCull *ctx, float *out6);

/* Tuning / reporting knobs (never change results unless stated). */
#define DPCU_CULL_OPT_KERNEL        1   /* DPCU_KERNEL_*: which exact form of the cull kernel runs           */
#define DPCU_KERNEL_AUTO    0           /* the measured winner: line-granular (list built in-kernel) for groups */
                                        /* of >= 6.7 M (1 view) / 7.0 M (2) / 3.6 M (3) / 3.0 M (4) / 2.6 M (5+ views) objects and when peer */
                                        /* bitsets are set; else direct                                         */
                                        /* (1 view) / view-sequential packed (>= 2 views)                       */
#define DPCU_KERNEL_DIRECT  1           /* one thread per object, scalar arithmetic, all views interleaved   */
#define DPCU_KERNEL_STAGED  2           /* persistent CTAs, TMA bulk + cp.async staging in shared memory     */
#define DPCU_KERNEL_VIEWS   3           /* views one after the other, packed f32x2 arithmetic                */
#define DPCU_KERNEL_LINES   4           /* a warp per 1024 objects, whole 128-byte bitset lines; always used */
                                        /* when peer bitsets are set (dpcuCullResultSetPeerBits)             */
#define DPCU_KERNEL_VIEWS_CHAINS 5      /* the views form with three predicate chains per axis instead of a  */
                                        /* counted compare (the earlier formulation; kept for comparison)    */
#define DPCU_KERNEL_FUSED_LEAF   6      /* reported by DPCU_CULL_OPT_LAST_KERNEL only (dpcuCullRunWithTree)  */
#define DPCU_KERNEL_LINES_PAIRS  7      /* >= 2 views: line-granular, two views per packed filter instruction, */
                                        /* undecided pairs queued across the line's steps, L2 bulk prefetch;   */
                                        /* what AUTO and peer bitsets use for several views (1 view: = LINES)  */
#define DPCU_KERNEL_GRID         8      /* one thread per object on a co-resident (cooperative) grid, the ordered */
                                        /* changed list behind one grid-wide barrier: one launch per cull for     */
                                        /* small groups (option; measured slower than direct + compaction)        */
#define DPCU_CULL_OPT_FMA           2   /* 1 = fused multiply-add fast mode: NOT bit-exact, reporting only   */
#define DPCU_CULL_OPT_CHANGED_LIST  3   /* 1 (default) = build the ordered changed list, 0 = bits only       */
#define DPCU_CULL_OPT_CTAS_PER_SM   4   /* 0 = auto                                                          */
#define DPCU_CULL_OPT_PROFILE       5   /* 1 = bracket every cull-kernel launch with CUDA events (see below) */
#define DPCU_CULL_OPT_FUSE_LEAF     6   /* 1 (default) = dpcuCullRunWithTree may run the tree's last level   */
                                        /* inside the cull kernel; 0 = always propagate, then cull           */
#define DPCU_CULL_OPT_FUSE_LIST     7   /* 1 (default) = the line-granular kernel builds the changed list itself   */
                                        /* (single-pass look-back); 0 = segment counters + compaction kernel       */
#define DPCU_CULL_OPT_FILTER        9   /* 1 (default) = multi-view culls decide provable (object, view) pairs from  */
                                        /* the OBB's centre and radius and run the reference arithmetic only on the  */
                                        /* rest (same bits, proof in cull_filter.cuh); 0 = reference arithmetic for all; */
                                        /* 2 / 3 = DIAGNOSTICS ONLY: margin shrunk to 1/8 (the error bound itself) / 0,  */
                                        /* used by the tests to measure the slack of the proof                           */
#define DPCU_CULL_OPT_LINE_WORDS   10   /* line-granular kernel: bitset words per warp; 0 (default) = 32 (a 128-byte */
                                        /* line); 8 / 16 = shorter lines (more warps on small groups; experiment)    */
#define DPCU_CULL_OPT_LIST_OFFSETS 11   /* one thread per object forms (direct, views, fused leaf): who turns flipped bits into list offsets */
                                        /* 0 (default) = AUTO; 1 = the cull kernel counts flips per 8192-object segment and its   */
                                        /* last CTA scans the counters; 2 = no counters at all, every compaction CTA popcounts    */
                                        /* the flipped-bit words before its segment (groups <= 2 Mi objects); 3 = counters, every */
                                        /* compaction CTA sums the counters before its segment (groups <= 32 Mi objects; no fence, */
                                        /* ticket or serial scan at the end of the cull kernel).  Larger groups fall back to 1     */
#define DPCU_CULL_OPT_L2_PREFETCH  12   /* 1 (default) = the line-granular kernels (1 view; 2-4 views) ask for every sector of the   */
                                        /* next step's matrices and extents a step ahead (prefetch.global.L2): 3 % faster per cull  */
                                        /* (64 Mi objects: 0.985 vs 1.014 ms; without the prefetch, index look-ahead only: 0.998 ms). */
                                        /* 0 = for callers that cull back to back for hundreds of milliseconds: under the board's     */
                                        /* power cap the extra L2 lookups cost more than they give (median of 300 culls: 1.059 ms    */
                                        /* with, 1.026 ms without, 1.049 ms for the round-1 kernel)                                   */
#define DPCU_CULL_OPT_LAST_KERNEL   8   /* read-only: DPCU_KERNEL_* form the last cull ran                          */
int dpcuCullSetOption(dpcuCull *ctx, int option, int value);
int dpcuCullGetOption(const dpcuCull *ctx, int option, int *value);
/* With DPCU_CULL_OPT_PROFILE = 1: device time spent in the cull kernel (K2 only, not the memset /
 * compaction around it) since the last call, measured with CUDA events on the launching stream;
 * synchronises those events and resets the accumulator.  This is the number bench.py's roofline uses. */
int dpcuCullGetKernelTime(dpcuCull *ctx, double *totalMs, uint64_t *launches);
/* the same per launch (oldest first, up to `capacity` entries; *launches = how many there were); also resets */
int dpcuCullGetKernelTimes(dpcuCull *ctx, float *perLaunchMs, size_t capacity, size_t *launches);
/* Diagnostics for tests/test_sass.py (no device needed): byte offsets, inside the cull kernels' parameter block
 * for nViews views, of the (1.0f, 1.0f) multiplier, the view-projection rows and the filter constants, so that the
 * SASS check can tell the reference arithmetic (products of view-projection entries must never be contracted into
 * a fused multiply-add) from the filter's own arithmetic (any rounding is fine there). */
int dpcuDebugKernelArgLayout(int nViews, size_t *onePairOffset, size_t *viewProjectionOffset, size_t *filterOffset,
                             size_t *totalBytes);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
int dpcuCullGetLaunchCount(const dpcuCull *ctx, uint64_t *launches);

/* ===================================================================== multi-GPU
 * Objects shard as contiguous slices: each GPU owns one dpcuCull over its slice.
 * When the consumer needs the full bitset on every GPU the cull kernel itself stores each
 * finished word into every peer's buffer over NVLink (peer-mapped stores) - the all-gather is
 * the kernel's epilogue, no separate collective is launched.
 *   peerBits[p] : device pointer (valid on ctx's device: cudaIpc-opened or peer-enabled) to
 *                 rank p's FULL bitset of this view, nPeers entries; entry == NULL is skipped
 *                 (use NULL for the local rank if the local result buffer is not the full one).
 *   wordOffset  : word index of this shard's first object inside the full bitset
 *                 (shard starts are multiples of 1024 objects). */
int dpcuCullResultSetPeerBits(dpcuCullResult *result, uint32_t *const *peerBits, int nPeers, size_t wordOffset);
/* cudaIpc plumbing so one-process-per-GPU launchers (torchrun) can exchange buffers */
#define DPCU_IPC_HANDLE_BYTES 64
int dpcuIpcGetHandle(const void *devicePointer, unsigned char handle[DPCU_IPC_HANDLE_BYTES]);
int dpcuIpcOpen(const unsigned char handle[DPCU_IPC_HANDLE_BYTES], void **devicePointer);
int dpcuIpcClose(void *devicePointer);
/* single-process multi-device launchers: enable direct access device -> peer */
int dpcuDeviceEnablePeerAccess(int device, int peer);

/* ===================================================================== transform layer
 * Device mirror of dp::transform::Tree (dp/transform/Tree.h:40-131): local and world matrices
 * resident in HBM, level-sorted {parent, transform} lists, dirty bit arrays; compute()
 * propagates world = local * world[parent] level by level for dirty nodes
 * (dp/transform/src/Tree.cpp:133-166). */
int dpcuTreeCreate(dpcuTree **out, int device);
int dpcuTreeDestroy(dpcuTree *tree);
/* Topology as Tree keeps it (Tree.h:113-128): entries = {parent, transform} u32 pairs of all
 * levels back to back; level l = entries[levelOffsets[l] .. levelOffsets[l+1]).  numNodes counts
 * index 0, the virtual identity root (Tree.cpp:42-47).  Replaces repeated addTransform /
 * removeTransform (Tree.cpp:54-98): the host keeps the index allocator, the device gets the lists.
 * New nodes start with identity local/world and dirty local bit set (Tree.cpp:63). */
int dpcuTreeSetTopology(dpcuTree *tree, const uint32_t *entries, const uint32_t *levelOffsets,
                        int numLevels, size_t numNodes);
/* updateLocalMatrix (Tree.h:85) for a contiguous range / a scattered batch; marks them dirty */
int dpcuTreeSetLocals(dpcuTree *tree, size_t first, size_t count, const float *matrices, int memspace);
int dpcuTreeUpdateLocals(dpcuTree *tree, const uint32_t *indices, size_t n, const float *matrices, int memspace);
/* mark [first, first+count) dirty without uploading (locals were written on the device) */
int dpcuTreeMarkDirty(dpcuTree *tree, size_t first, size_t count);
/* Tree::compute (Tree.cpp:133-166).  Afterwards the dirty-world set of this compute is
 * readable (the EventWorldMatricesChanged payload, Tree.h:46-58) until the next compute. */
int dpcuTreeCompute(dpcuTree *tree, dpcuStream *stream);
/* getWorldMatrices (Tree.h:82) - device pointer for dpcuCullBindMatrices, and host copies */
int dpcuTreeWorldDevicePointer(dpcuTree *tree, const float **deviceMatrices, size_t *numNodes);
int dpcuTreeLocalDevicePointer(dpcuTree *tree, float **deviceMatrices, size_t *numNodes);
int dpcuTreeGetWorld(dpcuTree *tree, size_t first, size_t count, float *hostMatrices);
int dpcuTreeGetDirtyWorld(dpcuTree *tree, uint32_t *hostWords, size_t nWords);
/* Refresh a host copy of the world matrices (hostWorld: numNodes x 16 floats, 64-byte stride, e.g.
 * dp::transform::Tree::m_matricesWorld) for exactly the nodes the last compute changed: the device compacts
 * them (index + matrix), one transfer brings them over, the host scatters them.  *updated = how many. */
int dpcuTreeGetWorldDirty(dpcuTree *tree, float *hostWorld, size_t numNodes, size_t *updated);
int dpcuTreeGetLaunchCount(const dpcuTree *tree, uint64_t *launches);
/* Tuning knob (never changes results): levels with at least this many nodes are propagated by the
 * persistent one-thread-per-node kernel with coalesced matrix traffic, smaller ones by the
 * four-threads-per-node kernel.  Default 65536; 0 = always the wide form, SIZE_MAX = never. */
#define DPCU_TREE_OPT_WIDE_MIN_NODES 1
int dpcuTreeSetOption(dpcuTree *tree, int option, size_t value);

/* One frame of the reference's update + cull (SceneTree::update -> Tree::compute, then
 * CullingImpl::cull: dp/sg/xbar/src/SceneTree.cpp:153-170, dp/sg/xbar/culling/src/CullingImpl.cpp:126-144)
 * in one call: dpcuTreeCompute followed by dpcuCullRun over the tree's world matrices (bound in
 * place, like dpcuCullBindMatrices) - same results, same dirty-set protocol.  When object i of
 * the context is bound to the transform of entry i of the tree's LAST level (one drawable per
 * leaf transform, in tree order; verified on the device and cached), that level is propagated
 * inside the cull kernel: the leaf world matrices are written out for the renderer and consumed
 * from shared memory, so they are not read back from HBM (-64 B per object).  Any other binding
 * runs the two steps back to back. */
int dpcuCullRunWithTree(dpcuCull *ctx, dpcuTree *tree, dpcuCullResult *const *results,
                        const float *viewProjections, int nViews, dpcuStream *stream);

/* ===================================================================== synthetic scenes (bench only)
 * On-device replay of pipeline_b200/scenes.py::random_objects: fills packed
 * lower4 (w = global object index - indexBase as u32 bits), extent4 and matrices for objects
 * [first, first+count).  Bit-identical to the host generator (tests/test_scene_gen.py). */
int dpcuSceneGenerate(uint64_t seed, uint64_t first, size_t count, uint32_t indexBase,
                      float *lower4Device, float *extent4Device, float *matricesDevice, dpcuStream *stream);
/* Bench support: stream `bytes` of device memory through L2 once (reads only).  bench.py runs it after its
 * L2-flushing write so that the L2 is cold and CLEAN when a timed step of a small workload starts. */
int dpcuDebugReadSweep(const void *deviceMemory, size_t bytes, dpcuStream *stream);

#ifdef __cplusplus
}
#endif
#endif /* DPCU_H */
