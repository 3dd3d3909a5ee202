// mex_stub.hpp — what a C++ test harness needs beyond mex.h when it links stub/mex_stub.cpp
#pragma once

#include <string>

// thrown by mexErrMsgTxt / mexErrMsgIdAndTxt.  Deliberately NOT derived from std::exception: like MATLAB's own
// error unwinding it must pass through the gateway's `catch (const std::exception &)` untouched.
struct MexError {
    std::string id, message;
    MexError(const std::string &id_, const std::string &msg) : id(id_), message(msg) {}
    const char *what() const { return message.c_str(); }
};

extern int mex_stub_lock_count;           // mexLock() minus mexUnlock()
extern std::string mex_stub_last_warning; // text of the last mexWarnMsgTxt
