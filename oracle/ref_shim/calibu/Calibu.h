// ORACLE — TEST INFRASTRUCTURE ONLY.  The reference's types.h only names calibu::Rig inside a member template that the
// hot path never instantiates (PoseT::GetTsw, types.h:73-85).
#pragma once
#include <memory>
#include <vector>
namespace calibu {
template <class Scalar> struct Rig;
}
