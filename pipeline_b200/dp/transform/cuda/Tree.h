// dp::transform::cuda::Tree - the transform hierarchy of dp::transform::Tree with compute() on the GPU.
//
// Derives the reference class (dp/transform/Tree.h:42-131) and overrides its one virtual hot
// function, compute() (dp/transform/src/Tree.cpp:133-166): the level-sorted {parent, transform}
// lists and the local matrices are mirrored into a dpcuTree (include/dpcu.h), dirty locals are
// pushed as one scattered batch per frame, the levels are propagated by the device kernels, and
// the published dirty-world set (EventWorldMatricesChanged, Tree.h:46-58) is the device's.
// The world matrices stay resident in HBM for the culler (dp::culling::cuda::Manager::
// groupSetDeviceMatrices - no re-upload); the host copy behind
// getWorldMatrices() is refreshed for the nodes that changed unless setHostWorldMirror(false).
//
// Drop-in: dp::sg::xbar::TransformTree holds its tree by value (dp/sg/xbar/TransformTree.h:89);
// changing that member's type to dp::transform::cuda::Tree is the whole patch.
#pragma once

#include <dp/transform/Tree.h>
#include <dpcu.h>

#include <vector>

namespace dp
{
  namespace transform
  {
    namespace cuda
    {

      class Tree : public dp::transform::Tree
      {
      public:
        explicit Tree( int device = 0 );
        virtual ~Tree();

        // addTransform / removeTransform are not virtual in the reference; these shadow them to note
        // that the level lists changed.  Callers that edit the tree through a base-class reference
        // call topologyChanged() themselves (compute() also notices changed level sizes).
        Index addTransform( Index parentIndex, dp::math::Mat44f const & matrix );
        void  removeTransform( Index transformIndex );
        void  topologyChanged() { m_topologyDirty = true; }

        //! \brief Tree::compute on the device; same dirty protocol, same notification, same results bit for bit
        virtual void compute( dp::math::Mat44f const & camera );

        //! \brief world matrices in HBM (64-byte stride, getTransformCount() of them) for groupSetDeviceMatrices
        void const * getDeviceWorldMatrices() const;
        dpcuTree *   getDeviceTree() const { return m_tree; }

        //! \brief keep m_matricesWorld (getWorldMatrices / getWorldMatrix) current after compute(); default true
        void setHostWorldMirror( bool enable ) { m_hostWorldMirror = enable; }

      private:
        Tree( Tree const & );
        Tree & operator=( Tree const & );

        void syncTopology();

        dpcuTree *            m_tree;
        bool                  m_topologyDirty;
        bool                  m_hostWorldMirror;
        std::vector<size_t>   m_levelSizes;       // level sizes of the topology the device holds
        std::vector<uint32_t> m_entries;          // staging: {parent, transform} pairs, levels back to back
        std::vector<uint32_t> m_levelOffsets;
        std::vector<uint32_t> m_indices;          // staging: dirty local indices
        std::vector<float>    m_matrices;         // staging: their matrices
        std::vector<uint32_t> m_words;            // staging: dirty-world words
      };

    } // namespace cuda
  } // namespace transform
} // namespace dp
