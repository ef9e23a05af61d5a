// pqb200_pic_calculators.h -- the reference's calculator classes, under the reference's own names in
// namespace pic, with their bodies replaced by calls into libpqb200.so (include/pqb200.h through the
// templates of include/pqb200_calculators.hpp).
//
// This is the patch a piquassoboost maintainer applies: put this directory IN FRONT of
// piquassoboost/sampling/source on the include path and the translation units that use the
// calculators -- the CPython wrappers sampling/GlynnPermanentCalculator_Wrapper.hpp:132-378 and
// sampling/ChinHuhPermanentCalculator_Wrapper.hpp:143-408, compiled UNMODIFIED -- construct these
// classes instead of the TBB ones.  Everything else stays the reference's: pic::matrix,
// pic::PicState_int64, pic::Complex16 (common/source/matrix.h, PicState.h) and numpy_interface.
// oracle/Makefile (`make dropin`) does exactly that and tests/test_reference_wrappers_gpu.py runs the
// reference's own unittests/test_permanent.py through the result.
//
//   reference class                                            declared in (piquassoboost/sampling/source/)
//   pic::GlynnPermanentCalculatorDouble / ...LongDouble        GlynnPermanentCalculator.h
//   pic::GlynnPermanentCalculatorSimpleDouble / ...LongDouble  GlynnPermanentCalculatorSimple.h
//   pic::BBFGPermanentCalculator                               BBFGPermanentCalculator.h:45-75
//   pic::CChinHuhPermanentCalculator                           CChinHuhPermanentCalculator.h:44-74
//   pic::GlynnPermanentCalculatorRepeatedDouble / ...LongDouble GlynnPermanentCalculatorRepeated.h:32-62
//   pic::BBFGPermanentCalculatorRepeated                       BBFGPermanentCalculatorRepeated.h
#ifndef PQB200_PIC_CALCULATORS_H
#define PQB200_PIC_CALCULATORS_H

#include "PicState.h"  // the reference's own containers (kept, not re-implemented)
#include "matrix.h"

#define PQB200_COMPLEX16 pic::Complex16
#include "../pqb200_calculators.hpp"

namespace pic {

class BBFGPermanentCalculator {
public:
    BBFGPermanentCalculator() {}
    /// BBFGPermanentCalculator.cpp:89-126; use_inf (MPFR there) is served by the double-double mode
    Complex16 calculate(matrix& mtx, bool use_extended = true, bool use_inf = false) {
        return pqb200::BBFGPermanentCalculator().calculate(mtx, use_extended, use_inf);
    }
};

class GlynnPermanentCalculatorDouble {
public:
    Complex16 calculate(matrix& mtx) { return pqb200::GlynnPermanentCalculator().calculate(mtx); }
};
class GlynnPermanentCalculatorLongDouble {
public:
    Complex16 calculate(matrix& mtx) { return pqb200::GlynnPermanentCalculator().calculate(mtx); }
};

class GlynnPermanentCalculatorSimpleDouble {
public:
    Complex16 calculate(matrix& mtx) { return pqb200::GlynnPermanentCalculatorSimple().calculate(mtx); }
};
class GlynnPermanentCalculatorSimpleLongDouble {
public:
    Complex16 calculate(matrix& mtx) { return pqb200::GlynnPermanentCalculatorSimple().calculate(mtx); }
};

class BBFGPermanentCalculatorRepeated {
public:
    BBFGPermanentCalculatorRepeated() {}
    /// BBFGPermanentCalculatorRepeated.cpp:74-141: col_mult = input state, row_mult = output state
    Complex16 calculate(matrix& mtx, PicState_int64& col_mult, PicState_int64& row_mult, bool use_extended = true,
                        bool use_inf = false) {
        if (mtx.rows == 0) return Complex16(1.0, 0.0);
        return pqb200::BBFGPermanentCalculatorRepeated().calculate(mtx, col_mult, row_mult, use_extended, use_inf);
    }
};

class GlynnPermanentCalculatorRepeatedDouble {
public:
    Complex16 calculate(matrix& mtx, PicState_int64& input_state, PicState_int64& output_state) {
        return pqb200::GlynnPermanentCalculatorRepeated().calculate(mtx, input_state, output_state);
    }
};
class GlynnPermanentCalculatorRepeatedLongDouble {
public:
    Complex16 calculate(matrix& mtx, PicState_int64& input_state, PicState_int64& output_state) {
        return pqb200::GlynnPermanentCalculatorRepeated().calculate(mtx, input_state, output_state);
    }
};

/// The reference's Ryser-type calculator (CChinHuhPermanentCalculator.cpp): the same permanent, so
/// the same engine serves it.
class CChinHuhPermanentCalculator {
public:
    CChinHuhPermanentCalculator() {}
    Complex16 calculate(matrix& mtx, PicState_int64& input_state, PicState_int64& output_state) {
        return pqb200::BBFGPermanentCalculatorRepeated().calculate(mtx, input_state, output_state);
    }
};

}  // namespace pic

#endif
