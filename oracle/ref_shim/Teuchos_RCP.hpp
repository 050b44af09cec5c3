// TEST INFRASTRUCTURE ONLY (oracle): minimal Teuchos::RCP stand-in so utils.h's extern declarations parse.
#pragma once
#include <memory>
#include <vector>
#include <string>
#include <iostream>
#include <iomanip>
#include <cmath>
#include <cstring>
namespace Teuchos {
  template <class T> class RCP { T *p_; public: RCP() : p_(nullptr) {} T *get() const { return p_; } T &operator*() const { return *p_; } T *operator->() const { return p_; } };
  class Time { public: Time(const char * = "") {} };
}
