// Model<double> and every Float64 kernel instantiation
#include "oc_model_impl.h"

namespace oc {
template class Model<double>;
}
