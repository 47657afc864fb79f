/* empty stub: the culling path includes <boost/scoped_array.hpp> but uses no Boost symbol */
