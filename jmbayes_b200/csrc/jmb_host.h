// jmb_host.h - host-side helpers shared by the translation units of libjmb200.so
#pragma once
#include <string>

// records the message jmb_last_error() returns on this thread; always returns 1
int jmb_set_error(const std::string& message);
