// Model<float> and every Float32 kernel instantiation
#include "oc_model_impl.h"

namespace oc {
template class Model<float>;
}
