#!/usr/bin/env python
"""Generate patches/0001-culling-cuda-backend.patch: the reference-side hookup of the cuda backend.

Run in the build container (needs /root/reference, read-only).  The edits below are applied to copies of the
reference files (line endings preserved - several of them are CRLF) and `diff -u` writes the patch; nothing of the
reference is copied into this repository except the context lines of the hunks.

    python tools/make_patch.py          ->  patches/0001-culling-cuda-backend.patch

tests/test_patch.py applies the patch to a fresh copy of those files (patch --dry-run and for real) and
tests/cpp/Makefile compiles the PATCHED CullingImpl.cpp for the frame-loop test.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DP_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "patches", "0001-culling-cuda-backend.patch")

EDITS = {
    # ---- the switch (CullingImpl.cpp:31-32,52-63), the per-frame matrix feed (:126-144), the dormant observer (:212-219)
    "dp/sg/xbar/culling/src/CullingImpl.cpp": [
        ('''#include <dp/culling/cpu/Manager.h>
#include <dp/culling/opengl/Manager.h>
''', '''#include <dp/culling/cpu/Manager.h>
#include <dp/culling/opengl/Manager.h>
#if defined(DP_CULLING_CUDA)
#include <dp/culling/cuda/Manager.h>
#include <dp/transform/cuda/Tree.h>
#endif
'''),
        ('''          : m_sceneTree( sceneTree )
        {
          switch ( cullingMode )''', '''          : m_sceneTree( sceneTree )
          , m_cudaManager( nullptr )
        {
          switch ( cullingMode )'''),
        ('''            m_culling.reset(dp::culling::opengl::Manager::create());
            break;
          default:''', '''            m_culling.reset(dp::culling::opengl::Manager::create());
            break;
#if defined(DP_CULLING_CUDA)
          case dp::culling::Mode::CUDA:
            // throws std::runtime_error without a usable CUDA device: the backend itself has no CPU path
            m_cudaManager = dp::culling::cuda::Manager::create();
            m_culling.reset( m_cudaManager );
            break;
#endif
          default:'''),
        ('''          // and attach to SceneTree get update events
          m_sceneTree->attach( this );
        }

        CullingImpl::~CullingImpl()
        {
          m_cullingGroup.reset();
          m_sceneTree->detach( this );
        }
''', '''          // and attach to SceneTree get update events
          m_sceneTree->attach( this );

          // world matrices that change are reported matrix by matrix (groupMatrixChanged): a backend that keeps its
          // own copy of the matrices (cuda, opengl) re-reads only those instead of all of them
          m_transformObserver.reset( new TransformObserver( *this ) );
          m_sceneTree->getTransformTree().getTree().attach( m_transformObserver.get() );
        }

        CullingImpl::~CullingImpl()
        {
          m_sceneTree->getTransformTree().getTree().detach( m_transformObserver.get() );
          m_cullingGroup.reset();
          m_sceneTree->detach( this );
        }

        void CullingImpl::setMatrices()
        {
          dp::transform::Tree & tree = m_sceneTree->getTransformTree().getTree();
#if defined(DP_CULLING_CUDA)
          // device-resident feed: the cuda transform tree's world matrices are culled in place, nothing is uploaded
          dp::transform::cuda::Tree * cudaTree = m_cudaManager ? dynamic_cast<dp::transform::cuda::Tree*>( &tree ) : nullptr;
          if ( cudaTree )
          {
            m_cudaManager->groupSetDeviceMatrices( m_cullingGroup, cudaTree->getDeviceWorldMatrices(), tree.getTransformCount() );
            return;
          }
#endif
          dp::math::Mat44f const * transforms = tree.getWorldMatrices();
          m_culling->groupSetMatrices(m_cullingGroup, transforms, tree.getTransformCount(), sizeof(transforms[0]));
        }
'''),
        ('''          ResultImplSharedPtr resultImpl = std::static_pointer_cast<ResultImpl>(result);
          dp::math::Mat44f const * transforms = m_sceneTree->getTransformTree().getTree().getWorldMatrices();
          m_culling->groupSetMatrices(m_cullingGroup, transforms, m_sceneTree->getTransformTree().getTree().getTransformCount(), sizeof(transforms[0]));
          m_culling->cull(''', '''          ResultImplSharedPtr resultImpl = std::static_pointer_cast<ResultImpl>(result);
          setMatrices();
          m_culling->cull('''),
        ('''        dp::math::Box3f CullingImpl::getBoundingBox()
        {
          dp::math::Mat44f const * transforms = m_sceneTree->getTransformTree().getTree().getWorldMatrices();
          m_culling->groupSetMatrices(m_cullingGroup, transforms, m_sceneTree->getTransformTree().getTree().getTransformCount(), sizeof(transforms[0]));
          return''', '''        dp::math::Box3f CullingImpl::getBoundingBox()
        {
          setMatrices();
          return'''),
    ],
    "dp/sg/xbar/culling/inc/CullingImpl.h": [
        ('''#include <dp/culling/Manager.h>

namespace dp
{
  namespace sg''', '''#include <dp/culling/Manager.h>

#include <memory>

namespace dp
{
  namespace culling
  {
    namespace cuda
    {
      class Manager;
    }
  }

  namespace sg'''),
        ('''          void updateBoundingBox( ObjectTreeIndex objectTreeIndex );
''', '''          void updateBoundingBox( ObjectTreeIndex objectTreeIndex );

          //! \\brief Hand the current world matrices of the transform tree to the culling group
          void setMatrices();
'''),
        ('''          std::unique_ptr<dp::culling::Manager>  m_culling;
''', '''          std::unique_ptr<dp::culling::Manager>  m_culling;
          dp::culling::cuda::Manager *              m_cudaManager;        // == m_culling.get() for Mode::CUDA, else nullptr
          std::unique_ptr<TransformObserver>        m_transformObserver;  // dirty world matrices -> groupMatrixChanged
'''),
    ],
    # ---- AUTO passed the member instead of the resolved mode (DrawableManagerDefault.cpp:529-535)
    "dp/sg/renderer/rix/gl/src/DrawableManagerDefault.cpp": [
        ('''Culling::create( getSceneTree(), m_cullingMode );''', '''Culling::create( getSceneTree(), cullingMode );'''),
    ],
    # ---- the transform tree member (TransformTree.h:89)
    "dp/sg/xbar/TransformTree.h": [
        ('''#include <dp/transform/Tree.h>
''', '''#include <dp/transform/Tree.h>
#if defined(DP_TRANSFORM_CUDA)
#include <dp/transform/cuda/Tree.h>
#endif
'''),
        ('''        dp::transform::Tree m_tree;
''', '''#if defined(DP_TRANSFORM_CUDA)
        dp::transform::cuda::Tree m_tree;   // compute() on the GPU, world matrices resident in HBM (throws without a CUDA device)
#else
        dp::transform::Tree m_tree;
#endif
'''),
    ],
    # ---- build: -DDPCU_HOME=<this repository> adds the backend
    "dp/culling/CMakeLists.txt": [
        ('''add_subdirectory( cpu )
add_subdirectory( opengl )
''', '''add_subdirectory( cpu )
add_subdirectory( opengl )

# dp::culling::cuda::Manager (B200): host layer from ${DPCU_HOME}/pipeline_b200/dp/culling/cuda over the C ABI of libdpcu.so
if ( DPCU_HOME )
  add_library( DPCullingCUDA STATIC
    "${DPCU_HOME}/pipeline_b200/dp/culling/cuda/Manager.h"
    "${DPCU_HOME}/pipeline_b200/dp/culling/cuda/inc/ManagerImpl.h"
    "${DPCU_HOME}/pipeline_b200/dp/culling/cuda/src/ManagerImpl.cpp"
  )
  target_include_directories( DPCullingCUDA PUBLIC "${DPCU_HOME}/pipeline_b200" "${DPCU_HOME}/include" )
  find_library( DPCU_LIBRARY dpcu PATHS "${DPCU_HOME}/pipeline_b200/lib" NO_DEFAULT_PATH )
  target_link_libraries( DPCullingCUDA DPUtil DPMath ${DPCU_LIBRARY} )
  set_target_properties( DPCullingCUDA PROPERTIES FOLDER "DP/Culling" )
  if(UNIX)
    set_target_properties( DPCullingCUDA PROPERTIES COMPILE_FLAGS -fPIC )
  endif()
endif()
'''),
        ('''target_link_libraries( DPCulling DPCullingCPU DPCullingOpenGL )
''', '''target_link_libraries( DPCulling DPCullingCPU DPCullingOpenGL )
if ( DPCU_HOME )
  target_link_libraries( DPCulling DPCullingCUDA )
endif()
'''),
    ],
    "dp/transform/CMakeLists.txt": [
        ('''target_link_libraries(DPTransform DPMath)
''', '''target_link_libraries(DPTransform DPMath)

# dp::transform::cuda::Tree (B200): Tree::compute on the device, world matrices stay in HBM for the culler
if ( DPCU_HOME )
  add_library( DPTransformCUDA STATIC
    "${DPCU_HOME}/pipeline_b200/dp/transform/cuda/Tree.h"
    "${DPCU_HOME}/pipeline_b200/dp/transform/cuda/src/Tree.cpp"
  )
  target_include_directories( DPTransformCUDA PUBLIC "${DPCU_HOME}/pipeline_b200" "${DPCU_HOME}/include" )
  find_library( DPCU_LIBRARY dpcu PATHS "${DPCU_HOME}/pipeline_b200/lib" NO_DEFAULT_PATH )
  target_link_libraries( DPTransformCUDA DPTransform DPUtil DPMath ${DPCU_LIBRARY} )
  set_target_properties( DPTransformCUDA PROPERTIES FOLDER "DP" )
  if(UNIX)
    set_target_properties( DPTransformCUDA PROPERTIES COMPILE_FLAGS -fPIC )
  endif()
endif()
'''),
    ],
    "dp/sg/xbar/culling/CMakeLists.txt": [
        ('''  DPCullingCPU
  DPCullingOpenGL
)
''', '''  DPCullingCPU
  DPCullingOpenGL
)

# dp::culling::cuda (B200 backend): sources and libdpcu.so come from the backend's tree, -DDPCU_HOME=<path to it>
if ( DPCU_HOME )
  target_link_libraries( DPSgXbarCulling DPCullingCUDA DPTransformCUDA )
  target_include_directories( DPSgXbarCulling PRIVATE "${DPCU_HOME}/pipeline_b200" "${DPCU_HOME}/include" )
  set_property( TARGET DPSgXbarCulling APPEND PROPERTY COMPILE_DEFINITIONS DP_CULLING_CUDA )
endif()
'''),
    ],
}


def main():
    if not os.path.isdir(REF):
        sys.exit("%s not found: the patch is generated where the reference tree exists" % REF)
    tmp = tempfile.mkdtemp(prefix="dpcu_patch_")
    try:
        for rel, edits in sorted(EDITS.items()):
            raw = open(os.path.join(REF, rel), "rb").read()
            crlf = b"\r\n" in raw
            text = raw.decode("latin-1")
            for old, new in edits:
                if crlf:
                    old, new = old.replace("\n", "\r\n"), new.replace("\n", "\r\n")
                assert text.count(old) == 1, "%s: context not found exactly once: %r" % (rel, old[:70])
                text = text.replace(old, new)
            for side, data in (("a", raw), ("b", text.encode("latin-1"))):
                path = os.path.join(tmp, side, rel)
                os.makedirs(os.path.dirname(path), exist_ok=True)
                open(path, "wb").write(data)
        res = subprocess.run(["diff", "-urN", "--label", "", "a", "b"], cwd=tmp, capture_output=True)
        # per-file labels instead of timestamps: run diff file by file
        out = []
        for rel in sorted(EDITS):
            r = subprocess.run(["diff", "-u", "--label", "a/" + rel, "--label", "b/" + rel, os.path.join("a", rel), os.path.join("b", rel)],
                               cwd=tmp, capture_output=True)
            assert r.returncode == 1, (rel, r.stderr)
            out.append(b"diff -u a/%s b/%s\n" % (rel.encode(), rel.encode()) + r.stdout)
        header = (b"Reference-side hookup of the dp::culling::cuda backend (nvpro-pipeline/pipeline).\n"
                  b"Apply from the root of the reference tree:  patch -p1 < 0001-culling-cuda-backend.patch\n"
                  b"Build with -DDPCU_HOME=<path to this backend's repository>; define DP_TRANSFORM_CUDA as well to make\n"
                  b"dp::sg::xbar::TransformTree compute on the GPU (its world matrices are then culled in place).\n"
                  b"Generated by tools/make_patch.py.\n\n")
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        open(OUT, "wb").write(header + b"".join(out))
        print("%s: %d bytes, %d files" % (OUT, os.path.getsize(OUT), len(EDITS)))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
