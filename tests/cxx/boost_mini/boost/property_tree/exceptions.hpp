// tests/cxx/boost_mini -- TEST INFRASTRUCTURE ONLY (see ptree.hpp): the exception classes live in ptree.hpp.
#ifndef DROPS_B200_TESTS_BOOST_MINI_EXCEPTIONS_HPP
#define DROPS_B200_TESTS_BOOST_MINI_EXCEPTIONS_HPP
#include <boost/property_tree/ptree.hpp>
#endif
