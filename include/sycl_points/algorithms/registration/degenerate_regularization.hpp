// DegenerateRegularizationParams / DegenerateRegularizationType — I/algorithms/registration/degenerate_regularization.hpp.
// The parameter structs live in registration_params.hpp here; the regularisation itself runs inside libspx
// (spx_degenerate_regularize, csrc/spx_registration.cu) between the device linearisation and the step.
#pragma once

#include "sycl_points/algorithms/registration/registration_params.hpp"
