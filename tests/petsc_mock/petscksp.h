#include "petsc_mock.h"
