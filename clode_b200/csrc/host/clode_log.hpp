// clode_log.hpp — small leveled logger with a replaceable sink.
// Stands in for the reference's use of spdlog (levels and call sites as in clode/cpp/CLODE.cpp,
// sinks as in clode/cpp/logging/PythonSink.hpp:22-37 and MatlabSink.hpp:22-41): the Python module
// installs a sink that prints through pybind11::print, everything else writes to stdout.
#pragma once

#include <cstdio>
#include <functional>
#include <mutex>
#include <sstream>
#include <string>

namespace clode_log {

enum level_enum { trace = 0, debug = 1, info = 2, warn = 3, err = 4, critical = 5, off = 6 };

struct Logger {
    level_enum level = info;
    std::function<void(const std::string &)> sink;
    std::string pattern;
    std::mutex mutex;

    static Logger &get()
    {
        static Logger lg;
        return lg;
    }

    void log(level_enum lvl, const std::string &msg)
    {
        if (lvl < level || level == off) return;
        static const char *names[] = {"trace", "debug", "info", "warning", "error", "critical", "off"};
        std::string line = std::string("[") + names[lvl] + "] " + msg;
        std::lock_guard<std::mutex> lock(mutex);
        if (sink) sink(line);
        else std::printf("%s\n", line.c_str());
    }
};

// "{}" placeholders, like the fmt strings at the reference call sites
inline void format_into(std::ostringstream &os, const char *fmt)
{
    os << fmt;
}
template <typename T, typename... Rest>
void format_into(std::ostringstream &os, const char *fmt, const T &value, const Rest &...rest)
{
    for (; *fmt; ++fmt) {
        if (fmt[0] == '{' && fmt[1] == '}') {
            os << value;
            format_into(os, fmt + 2, rest...);
            return;
        }
        os << *fmt;
    }
}
template <typename... Args> std::string format(const char *fmt, const Args &...args)
{
    std::ostringstream os;
    format_into(os, fmt, args...);
    return os.str();
}

template <typename... Args> void trace_(const char *f, const Args &...a) { Logger::get().log(trace, format(f, a...)); }
template <typename... Args> void debug_(const char *f, const Args &...a) { Logger::get().log(debug, format(f, a...)); }
template <typename... Args> void info_(const char *f, const Args &...a) { Logger::get().log(info, format(f, a...)); }
template <typename... Args> void warn_(const char *f, const Args &...a) { Logger::get().log(warn, format(f, a...)); }
template <typename... Args> void error_(const char *f, const Args &...a) { Logger::get().log(err, format(f, a...)); }

} // namespace clode_log
