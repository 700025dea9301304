/* MOCK of <caml/bigarray.h> */
#ifndef CAML_MOCK_BIGARRAY_H
#define CAML_MOCK_BIGARRAY_H
#include "mlvalues.h"
struct caml_ba_array { void *data; intnat num_dims; intnat flags; void *proxy; intnat dim[16]; };
struct caml_ba_array *caml_mock_ba_array_val(value v);
#define Caml_ba_array_val(v) caml_mock_ba_array_val(v)
#define Caml_ba_data_val(v) (Caml_ba_array_val(v)->data)
enum { CAML_BA_FLOAT64 = 1, CAML_BA_C_LAYOUT = 0, CAML_BA_EXTERNAL = 0, CAML_BA_MANAGED = 0x200 };
value caml_ba_alloc(int flags, int num_dims, void *data, intnat *dim);
#endif
