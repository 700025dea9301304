/* MOCK of <caml/alloc.h> */
#ifndef CAML_MOCK_ALLOC_H
#define CAML_MOCK_ALLOC_H
#include "mlvalues.h"
value caml_alloc_tuple(mlsize_t n);
value caml_copy_double(double d);
#endif
