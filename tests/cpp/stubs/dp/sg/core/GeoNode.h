// TEST INFRASTRUCTURE - minimal stand-in for the reference's dp/sg/core/GeoNode.h, just enough of the scene graph
// for the (patched) dp/sg/xbar/culling/src/CullingImpl.cpp to compile outside the renderer: CullingImpl only asks a
// GeoNode for its bounding box (CullingImpl.cpp:153-157).
#pragma once

#include <dp/math/Boxnt.h>

#include <memory>

namespace dp
{
  namespace sg
  {
    namespace core
    {
      class Object
      {
      public:
        virtual ~Object() {}
      };
      typedef std::shared_ptr<Object> ObjectSharedPtr;

      class GeoNode : public Object
      {
      public:
        explicit GeoNode( dp::math::Box3f const & box ) : m_box( box ) {}
        dp::math::Box3f const & getBoundingBox() const { return m_box; }
        void setBoundingBox( dp::math::Box3f const & box ) { m_box = box; }
      private:
        dp::math::Box3f m_box;
      };
      typedef std::shared_ptr<GeoNode> GeoNodeSharedPtr;
    }
  }
}
