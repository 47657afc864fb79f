// dp::culling::cuda::Manager - the B200 culling backend behind the unchanged dp::culling::Manager
// API.  Same shape as the reference's factories dp/culling/cpu/Manager.h:39-43 and
// dp/culling/opengl/Manager.h:38-42; the name is the one the reference's own (dead) code already
// expects: dp/sg/renderer/rix/common/src/DrawableManager.cpp:32,200-201.
//
// Select it in dp/sg/xbar/culling/src/CullingImpl.cpp:52-63 with
//     case dp::culling::Mode::CUDA: m_culling.reset( dp::culling::cuda::Manager::create() ); break;
// (INTEGRATION.md).  The implementation talks to the GPU only through the C ABI in include/dpcu.h.
#pragma once

#include <dp/culling/Config.h>
#include <dp/culling/ManagerBitSet.h>

namespace dp
{
  namespace culling
  {
    namespace cuda
    {

      class Manager : public dp::culling::ManagerBitSet
      {
      public:
        /** \brief Create the CUDA culling manager on the given device.
            \remarks throws std::runtime_error when no CUDA device is usable; there is no CPU fallback. **/
        DP_CULLING_API static Manager* create( int device = 0 );

        /** \brief Zero-copy matrix feed (extension): cull straight out of a device-resident matrix array,
                   e.g. the world matrices of a dpcuTree.  Replaces groupSetMatrices for this group until
                   groupSetMatrices is called again. 64-byte stride. **/
        DP_CULLING_API virtual void groupSetDeviceMatrices( GroupSharedPtr const & group, void const * deviceMatrices, size_t numberOfMatrices ) = 0;

        /** \brief Cull one group against several view-projections in one pass over the objects (extension).
                   results[v] receives the outcome for viewProjections[v]; at most 8 views per call. **/
        DP_CULLING_API virtual void cullMultiView( GroupSharedPtr const & group, std::vector<ResultSharedPtr> const & results
                                                 , std::vector<dp::math::Mat44f> const & viewProjections ) = 0;
      };

    } // namespace cuda
  } // namespace culling
} // namespace dp
