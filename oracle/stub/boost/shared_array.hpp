/* empty stub: the culling path includes <boost/shared_array.hpp> but uses no Boost symbol */
