// src/GluonFields.cpp is a missing large blob upstream (.MISSING_LARGE_BLOBS); these two
// arrays are only read when gluon_field_fluctuations=1 (Nucleus.cpp:34), which is out of scope.
// This stub exists only so the reference links.
#include "GluonField.h"
double GluonField::ImprintArrayLHC[600][600];
double GluonField::ImprintArrayRHIC[600][600];
