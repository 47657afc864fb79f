// dp::transform::cuda::Tree - host side; every GPU operation goes through the C ABI of include/dpcu.h.
#include <dp/transform/cuda/Tree.h>

#include <cstring>
#include <stdexcept>
#include <string>

namespace dp
{
  namespace transform
  {
    namespace cuda
    {
      namespace
      {
        inline void verify( int status, char const * call )
        {
          if ( status != DPCU_OK )
          {
            throw std::runtime_error( std::string( call ) + ": " + dpcuGetLastError() );
          }
        }
#define DPCU_VERIFY( call ) verify( call, #call )

        struct IndexCollector
        {
          IndexCollector( std::vector<uint32_t> & indices ) : m_indices( indices ) {}
          void operator()( size_t index ) { m_indices.push_back( static_cast<uint32_t>( index ) ); }
          std::vector<uint32_t> & m_indices;
        };
      }

      Tree::Tree( int device )
        : m_tree( nullptr )
        , m_topologyDirty( true )
        , m_hostWorldMirror( true )
      {
        DPCU_VERIFY( dpcuTreeCreate( &m_tree, device ) );   // throws without a GPU: there is no CPU path in this class
      }

      Tree::~Tree()
      {
        dpcuTreeDestroy( m_tree );
      }

      Index Tree::addTransform( Index parentIndex, dp::math::Mat44f const & matrix )
      {
        m_topologyDirty = true;
        return dp::transform::Tree::addTransform( parentIndex, matrix );
      }

      void Tree::removeTransform( Index transformIndex )
      {
        m_topologyDirty = true;
        dp::transform::Tree::removeTransform( transformIndex );
      }

      void const * Tree::getDeviceWorldMatrices() const
      {
        float const * world = nullptr;
        size_t count = 0;
        DPCU_VERIFY( dpcuTreeWorldDevicePointer( m_tree, &world, &count ) );
        return world;
      }

      void Tree::syncTopology()
      {
        // safety net for edits made through a base-class reference: a changed level size always shows
        if ( !m_topologyDirty )
        {
          m_topologyDirty = m_levelSizes.size() != m_transformLevels.size();
          for ( size_t level = 0; !m_topologyDirty && level < m_transformLevels.size(); ++level )
          {
            m_topologyDirty = m_levelSizes[level] != m_transformLevels[level].transformListEntries.size();
          }
        }
        if ( !m_topologyDirty )
        {
          return;
        }
        m_entries.clear();
        m_levelOffsets.assign( 1, 0 );
        m_levelSizes.clear();
        for ( size_t level = 0; level < m_transformLevels.size(); ++level )
        {
          TransformListEntries const & entries = m_transformLevels[level].transformListEntries;
          // TransformListEntry is two unsigned ints {parent, transform} (Tree.h:113-116): the device layout
          size_t const offset = m_entries.size();
          m_entries.resize( offset + 2 * entries.size() );
          if ( !entries.empty() )
          {
            memcpy( &m_entries[offset], entries.data(), entries.size() * sizeof(TransformListEntry) );
          }
          m_levelOffsets.push_back( static_cast<uint32_t>( m_entries.size() / 2 ) );
          m_levelSizes.push_back( entries.size() );
        }
        // the node arrays have the size of the host arrays (grown in 65536 steps, Tree.cpp:35,117-126);
        // existing local / world matrices stay where they are on the device
        DPCU_VERIFY( dpcuTreeSetTopology( m_tree, m_entries.data(), m_levelOffsets.data()
                                        , static_cast<int>( m_transformLevels.size() ), m_matricesLocal.size() ) );
        m_topologyDirty = false;
      }

      void Tree::compute( dp::math::Mat44f const & /*camera*/ )
      {
        static_assert( sizeof(TransformListEntry) == 2 * sizeof(uint32_t), "entry layout" );
        static_assert( sizeof(dp::math::Mat44f) == 16 * sizeof(float), "matrix layout" );

        syncTopology();

        // locals written since the last compute (updateLocalMatrix / addTransform set m_dirtyTransforms)
        m_indices.clear();
        IndexCollector collector( m_indices );
        m_dirtyTransforms.traverseBits( collector );
        if ( !m_indices.empty() )
        {
          size_t const first = m_indices.front(), count = m_indices.size();
          if ( m_indices.back() - first + 1 == count )
          {
            // one contiguous run (whole-scene animation): straight from the host array
            DPCU_VERIFY( dpcuTreeSetLocals( m_tree, first, count, m_matricesLocal[first].getPtr(), DPCU_MEM_HOST ) );
          }
          else
          {
            m_matrices.resize( 16 * count );
            for ( size_t i = 0; i < count; ++i )
            {
              memcpy( &m_matrices[16 * i], m_matricesLocal[m_indices[i]].getPtr(), 16 * sizeof(float) );
            }
            DPCU_VERIFY( dpcuTreeUpdateLocals( m_tree, m_indices.data(), count, m_matrices.data(), DPCU_MEM_HOST ) );
          }
        }

        DPCU_VERIFY( dpcuTreeCompute( m_tree, nullptr ) );

        // the set of world matrices this compute changed (Tree.cpp:155-158), from the device
        size_t const nodes = m_matricesLocal.size();
        m_words.resize( ( nodes + 31 ) / 32 + 2 );
        DPCU_VERIFY( dpcuTreeGetDirtyWorld( m_tree, m_words.data(), m_words.size() ) );
        m_dirtyWorldMatrices.setBits( m_words.data(), m_dirtyWorldMatrices.getSize() );

        if ( m_hostWorldMirror )
        {
          // refresh the host copy behind getWorldMatrices() for exactly the nodes that changed: the device compacts
          // them, one transfer brings them over (two dirty nodes at the ends of a 17.9 M-node tree are 136 bytes,
          // not the 1.15 GB span between them)
          size_t updated = 0;
          DPCU_VERIFY( dpcuTreeGetWorldDirty( m_tree, &m_matricesWorld[0][0][0], nodes, &updated ) );
        }

        notifyTransformsChanged( m_dirtyWorldMatrices );

        m_dirtyTransforms.clear();
        m_dirtyWorldMatrices.clear();
      }

    } // namespace cuda
  } // namespace transform
} // namespace dp
