// Internal view of the transform layer (dpcu_tree.cu) for the culling layer's fused
// propagate-and-cull path.  Not part of the C ABI.
#pragma once

#include "dpcu_internal.h"

#include <vector>

struct dpcuTree
{
  int          device = 0;
  cudaStream_t stream = nullptr;
  dpcu::DeviceArray local, world, entries, dirtyLocal, dirtyWorld, scratch;
  dpcu::DeviceArray gather;      // dpcuTreeGetWorldDirty: compacted {node index, world matrix} records
  dpcu::PinnedArray gatherHost;
  size_t       numNodes = 0;
  size_t       numEntries = 0;
  std::vector<uint32_t> levelOffsets;
  int          smCount = 1, wideCtasPerSm = 1;
  size_t       wideMinNodes = size_t( 1 ) << 16;   // levels at least this large run treeLevelWideKernel
  uint64_t     launches = 0;
  uint64_t     topologyVersion = 0;   // bumped by dpcuTreeSetTopology (cached leaf bindings of cull contexts go stale)
  dpcu::StreamFence done;        // last compute submitted
  dpcu::StreamFence readers;     // last cull that read `world` in place (dpcuCullBindTree / dpcuCullRunWithTree): the next compute waits for it
  dpcu::StreamFence uploads;
};

namespace dpcu
{
  int treeBeginCompute( dpcuTree *t, cudaStream_t s );
  int treeComputeLevels( dpcuTree *t, cudaStream_t s, size_t firstLevel, size_t lastLevel );
  int treeEndCompute( dpcuTree *t, cudaStream_t s );
}
