/* empty stub: the culling path includes <boost/assert.hpp> but uses no Boost symbol */
