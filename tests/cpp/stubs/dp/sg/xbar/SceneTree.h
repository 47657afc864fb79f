// TEST INFRASTRUCTURE - minimal stand-in for the reference's dp/sg/xbar/SceneTree.h (+ ObjectTree.h, Tree.h,
// TransformTree.h): the subset of the interface that dp/sg/xbar/culling/src/CullingImpl.cpp uses, so that the PATCHED
// CullingImpl (patches/0001-culling-cuda-backend.patch) can be compiled and driven without the scene graph, the
// effect system and the renderer behind the real SceneTree.  Names and signatures follow the reference:
//   ObjectTreeNode::m_transform / m_isDrawable / m_object      dp/sg/xbar/ObjectTree.h
//   PreOrderTreeTraverser<Tree, Visitor>::traverse              dp/sg/xbar/Tree.h
//   SceneTree::getObjectTree / getObjectTreeNode / getTransformTree / Event / attach   dp/sg/xbar/SceneTree.h
//   TransformTree::getTree                                      dp/sg/xbar/TransformTree.h:70
// The flat object "tree" here has no hierarchy of its own (every node is a root-level sibling); the transform
// hierarchy is the real dp::transform::Tree (or dp::transform::cuda::Tree).
#pragma once

#include <dp/sg/core/GeoNode.h>
#include <dp/transform/Tree.h>
#include <dp/util/Observer.h>
#include <dp/util/PointerTypes.h>

#include <cstdint>
#include <memory>
#include <vector>

namespace dp
{
  namespace sg
  {
    namespace xbar
    {
      typedef uint32_t ObjectTreeIndex;
      typedef uint32_t TransformIndex;

      struct ObjectTreeNode
      {
        ObjectTreeNode() : m_transform( 0 ), m_isDrawable( false ), m_alive( false ) {}
        TransformIndex                m_transform;
        bool                          m_isDrawable;
        dp::sg::core::ObjectSharedPtr m_object;
        bool                          m_alive;       // stub only: slot in use
      };

      class ObjectTree
      {
      public:
        size_t size() const { return m_nodes.size(); }
        ObjectTreeNode const & operator[]( size_t index ) const { return m_nodes[index]; }
        ObjectTreeNode & operator[]( size_t index ) { return m_nodes[index]; }
        std::vector<ObjectTreeNode> m_nodes;
      };

      template <typename TreeType, typename Visitor>
      class PreOrderTreeTraverser
      {
      public:
        void traverse( TreeType const & tree, Visitor & visitor )
        {
          typename Visitor::Data data;
          for ( size_t index = 0; index < tree.size(); ++index )
          {
            if ( tree[index].m_alive && visitor.preTraverse( static_cast<ObjectTreeIndex>( index ), data ) )
            {
              visitor.postTraverse( static_cast<ObjectTreeIndex>( index ), data );
            }
          }
        }
      };

      class TransformTree
      {
      public:
        explicit TransformTree( dp::transform::Tree & tree ) : m_tree( tree ) {}
        dp::transform::Tree & getTree() { return m_tree; }
      private:
        dp::transform::Tree & m_tree;
      };

      DEFINE_PTR_TYPES( SceneTree );

      class SceneTree : public dp::util::Subject
      {
      public:
        class Event : public dp::util::Event
        {
        public:
          enum class Type { ADDED, REMOVED, CHANGED, TRAVERSAL_MASK_CHANGED };
          Event( ObjectTreeIndex index, ObjectTreeNode const & node, Type type ) : m_index( index ), m_node( node ), m_type( type ) {}
          ObjectTreeIndex getIndex() const { return m_index; }
          ObjectTreeNode const & getNode() const { return m_node; }
          Type getType() const { return m_type; }
        private:
          ObjectTreeIndex        m_index;
          ObjectTreeNode const & m_node;
          Type                   m_type;
        };

        static SceneTreeSharedPtr create( dp::transform::Tree & tree ) { return SceneTreeSharedPtr( new SceneTree( tree ) ); }

        ObjectTree const & getObjectTree() const { return m_objectTree; }
        ObjectTreeNode const & getObjectTreeNode( ObjectTreeIndex index ) const { return m_objectTree[index]; }
        TransformTree & getTransformTree() { return m_transformTree; }

        // --- what the real SceneTree does when the scene graph changes (SceneTree.cpp: addDrawable / removeObjectTreeIndex /
        //     the GeoNode observer): edit the object tree, then notify the attached observers
        ObjectTreeIndex addDrawable( dp::math::Box3f const & box, TransformIndex transform )
        {
          ObjectTreeNode node;
          node.m_transform = transform;
          node.m_isDrawable = true;
          node.m_alive = true;
          node.m_object = std::make_shared<dp::sg::core::GeoNode>( box );
          m_objectTree.m_nodes.push_back( node );
          ObjectTreeIndex index = static_cast<ObjectTreeIndex>( m_objectTree.size() - 1 );
          notify( Event( index, m_objectTree[index], Event::Type::ADDED ) );
          return index;
        }
        void removeDrawable( ObjectTreeIndex index )
        {
          notify( Event( index, m_objectTree[index], Event::Type::REMOVED ) );
          m_objectTree[index].m_alive = false;
          m_objectTree[index].m_isDrawable = false;
          m_objectTree[index].m_object.reset();
        }
        void changeBoundingBox( ObjectTreeIndex index, dp::math::Box3f const & box )
        {
          std::static_pointer_cast<dp::sg::core::GeoNode>( m_objectTree[index].m_object )->setBoundingBox( box );
          notify( Event( index, m_objectTree[index], Event::Type::CHANGED ) );
        }

      private:
        explicit SceneTree( dp::transform::Tree & tree ) : m_transformTree( tree ) {}
        ObjectTree    m_objectTree;
        TransformTree m_transformTree;
      };
    }
  }
}
