// TEST INFRASTRUCTURE — stand-in for boost::format as lsq_registration_impl.hpp uses it (the LM debug table): the
// arguments are streamed one after the other, the format string is ignored.
#ifndef APDO_REF_STUB_BOOST_FORMAT
#define APDO_REF_STUB_BOOST_FORMAT
#include <ostream>
#include <sstream>
#include <string>
namespace boost {
class format {
 public:
  explicit format(const std::string&) {}
  template <class T>
  format& operator%(const T& v) {
    s_ << v << ' ';
    return *this;
  }
  std::string str() const { return s_.str(); }

 private:
  std::ostringstream s_;
};
inline std::ostream& operator<<(std::ostream& o, const format& f) { return o << f.str(); }
}  // namespace boost
#endif
