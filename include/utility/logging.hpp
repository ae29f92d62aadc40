// DEFAULT_SOURCE_LOG_* — the logging macro family the reference's drivers call
// (include/utility/logging.hpp:32-140 of the reference: TRACE < DEBUG < INFO < PROGRESS < WARNING <
// ERROR < FATAL, std::format-style arguments, compile-time level cut-off).  Minimal stand-in for the
// drop-in headers: one formatted line on std::clog, no logger objects, no spdlog.
// -DAMR_LOG_LEVEL=<TRACE|...|OFF> selects the cut-off like the reference's CMake option does;
// the default here is WARNING so that per-step progress lines do not serialise a GPU time loop.
#ifndef AMRB_UTILITY_LOGGING_HPP
#define AMRB_UTILITY_LOGGING_HPP
#include <format>
#include <iostream>
#include <string_view>

#define AMRB_LOG_RANK_TRACE 1
#define AMRB_LOG_RANK_DEBUG 2
#define AMRB_LOG_RANK_INFO 3
#define AMRB_LOG_RANK_PROGRESS 4
#define AMRB_LOG_RANK_WARNING 5
#define AMRB_LOG_RANK_ERROR 6
#define AMRB_LOG_RANK_FATAL 7
#define AMRB_LOG_RANK_OFF 8
#define AMRB_LOG_CAT_(a, b) a##b
#define AMRB_LOG_CAT(a, b) AMRB_LOG_CAT_(a, b)
#ifdef AMR_LOG_LEVEL
#    define AMRB_LOG_CUTOFF AMRB_LOG_CAT(AMRB_LOG_RANK_, AMR_LOG_LEVEL)
#else
#    define AMRB_LOG_CUTOFF AMRB_LOG_RANK_WARNING
#endif

namespace amr::utility::logging
{
template <typename... Args>
inline void emit(std::string_view tag, std::format_string<Args...> fmt, Args&&... args)
{
    std::clog << '[' << tag << "] " << std::format(fmt, std::forward<Args>(args)...) << '\n';
}
} // namespace amr::utility::logging

#define AMRB_LOG_AT(rank, tag, ...)                                                              \
    do                                                                                           \
    {                                                                                            \
        if constexpr ((rank) >= AMRB_LOG_CUTOFF) ::amr::utility::logging::emit(tag, __VA_ARGS__); \
    } while (0)

#define DEFAULT_SOURCE_LOG_TRACE(...) AMRB_LOG_AT(AMRB_LOG_RANK_TRACE, "trace", __VA_ARGS__)
#define DEFAULT_SOURCE_LOG_DEBUG(...) AMRB_LOG_AT(AMRB_LOG_RANK_DEBUG, "debug", __VA_ARGS__)
#define DEFAULT_SOURCE_LOG_INFO(...) AMRB_LOG_AT(AMRB_LOG_RANK_INFO, "info", __VA_ARGS__)
#define DEFAULT_SOURCE_LOG_PROGRESS(...) AMRB_LOG_AT(AMRB_LOG_RANK_PROGRESS, "progress", __VA_ARGS__)
#define DEFAULT_SOURCE_LOG_WARNING(...) AMRB_LOG_AT(AMRB_LOG_RANK_WARNING, "warning", __VA_ARGS__)
#define DEFAULT_SOURCE_LOG_ERROR(...) AMRB_LOG_AT(AMRB_LOG_RANK_ERROR, "error", __VA_ARGS__)
#define DEFAULT_SOURCE_LOG_FATAL(...) AMRB_LOG_AT(AMRB_LOG_RANK_FATAL, "fatal", __VA_ARGS__)
#endif
