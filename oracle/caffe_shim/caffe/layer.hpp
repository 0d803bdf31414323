// oracle/caffe_shim -- TEST INFRASTRUCTURE: forwards to the stand-in declarations in shim.hpp
#include "caffe/shim.hpp"
