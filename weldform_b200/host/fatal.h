// Host-side mirror of External/fatal.h + string_.h: `throw new Fatal(fmt, ...)` caught by MECHSYS_CATCH in main().
#ifndef WF_HOST_FATAL_H
#define WF_HOST_FATAL_H
#include <cstdarg>
#include <cstdio>
#include <exception>
#include <iostream>
#include <string>

class String : public std::string {
public:
    String() {}
    String(const std::string & s) : std::string(s) {}
    String(const char * s) : std::string(s) {}
    int Printf(const char * fmt, ...) { char b[2048]; va_list ap; va_start(ap, fmt); int n = vsnprintf(b, sizeof(b), fmt, ap); va_end(ap); assign(b); return n; }
    const char * CStr() const { return c_str(); }
};

class Fatal {
public:
    Fatal(const char * fmt, ...) { char b[2048]; va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof(b), fmt, ap); va_end(ap); msg_ = b; }
    void Cout() const { std::cout << "Fatal: " << msg_ << std::endl; }
    const std::string & Msg() const { return msg_; }
private:
    std::string msg_;
};

#define MECHSYS_CATCH \
    catch (Fatal * e) { e->Cout(); delete e; return 1; } \
    catch (char const * m) { std::cout << "Fatal: " << m << std::endl; return 1; } \
    catch (std::exception & e) { std::cout << "Fatal: " << e.what() << std::endl; return 1; } \
    catch (...) { std::cout << "Some exception (...) occurred" << std::endl; return 1; }

#endif
