// amr::cuda memory & sync wrappers with the reference's names and signatures
// (include/cuda/device_buffer.hpp:9-48), forwarding to the C ABI.
#ifndef AMR_INCLUDED_CUDA_DEVICE_BUFFER
#define AMR_INCLUDED_CUDA_DEVICE_BUFFER
#include "amrb_check.hpp"
#include <cstddef>

namespace amr::cuda
{
inline auto device_malloc(std::size_t bytes) -> void*
{
    void* p = nullptr;
    detail::check(amrb_device_malloc(&p, bytes), "device_malloc");
    return p;
}
inline auto device_free(void* ptr) noexcept -> void { amrb_device_free(ptr); }
inline auto host_pinned_malloc(std::size_t bytes) -> void*
{
    void* p = nullptr;
    detail::check(amrb_host_pinned_malloc(&p, bytes), "host_pinned_malloc");
    return p;
}
inline auto host_pinned_free(void* ptr) noexcept -> void { amrb_host_pinned_free(ptr); }
inline auto copy_host_to_device(void* dst, const void* src, std::size_t bytes) -> void
{
    detail::check(amrb_copy_host_to_device(dst, src, bytes), "copy_host_to_device");
}
inline auto copy_host_to_device_async(void* dst, const void* src, std::size_t bytes) -> void
{
    detail::check(amrb_copy_host_to_device_async(dst, src, bytes, nullptr), "copy_host_to_device_async");
}
inline auto copy_device_to_host(void* dst, const void* src, std::size_t bytes) -> void
{
    detail::check(amrb_copy_device_to_host(dst, src, bytes), "copy_device_to_host");
}
inline auto copy_device_to_host_async(void* dst, const void* src, std::size_t bytes) -> void
{
    detail::check(amrb_copy_device_to_host_async(dst, src, bytes, nullptr), "copy_device_to_host_async");
}
inline auto copy_device_to_host_async_on_stream(void* dst, const void* src, std::size_t bytes, void* stream) -> void
{
    detail::check(amrb_copy_device_to_host_async(dst, src, bytes, stream), "copy_device_to_host_async_on_stream");
}
inline auto copy_device_to_device(void* dst, const void* src, std::size_t bytes) -> void
{
    detail::check(amrb_copy_device_to_device(dst, src, bytes), "copy_device_to_device");
}
inline auto async_copy_stream_create() -> void*
{
    void* s = nullptr;
    detail::check(amrb_stream_create(&s), "async_copy_stream_create");
    return s;
}
inline auto async_copy_stream_destroy(void* stream) noexcept -> void { amrb_stream_destroy(stream); }
inline auto async_copy_stream_wait_for_fence(void* stream, void* fence) -> void
{
    detail::check(amrb_stream_wait_fence(stream, fence), "async_copy_stream_wait_for_fence");
}
inline auto async_copy_fence_create() -> void*
{
    void* f = nullptr;
    detail::check(amrb_fence_create(&f), "async_copy_fence_create");
    return f;
}
inline auto async_copy_fence_destroy(void* fence) noexcept -> void { amrb_fence_destroy(fence); }
inline auto async_copy_fence_record(void* fence) -> void
{
    detail::check(amrb_fence_record(fence, nullptr), "async_copy_fence_record");
}
inline auto async_copy_fence_record_on_stream(void* fence, void* stream) -> void
{
    detail::check(amrb_fence_record(fence, stream), "async_copy_fence_record_on_stream");
}
inline auto async_copy_fence_wait(void* fence) -> void
{
    detail::check(amrb_fence_wait(fence), "async_copy_fence_wait");
}
} // namespace amr::cuda
#endif
