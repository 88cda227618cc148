// host_error.h — error text for calls that have no context (thread-local slot).
#pragma once
namespace curvis {
int set_thread_error(int code, const char* msg);  // stores msg, returns code
const char* thread_error();
}
