// Errors.hpp -- std::runtime_error with stream-style message building (the reference uses bstl::throw_ex,
// common/bstl/include/bstl/throw_stdex.hpp:8-10; the message strings are pinned by its tests).
#pragma once
#include <sstream>
#include <stdexcept>
#include <string>

namespace bl::llama {

class Raise {
public:
    Raise() = default;
    [[noreturn]] ~Raise() noexcept(false) { throw std::runtime_error(m_text.str()); }
    template <class T> Raise& operator<<(const T& v) { m_text << v; return *this; }
private:
    std::ostringstream m_text;
};

// converts a failed C-ABI status into an exception carrying blk_last_error()
void throwIfFailed(int status, const char* what);

} // namespace bl::llama
