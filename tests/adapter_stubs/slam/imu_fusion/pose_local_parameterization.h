// stub: see ../stub_reference.h
#pragma once
#include "stub_reference.h"
