// tests/cxx/boost_mini -- TEST INFRASTRUCTURE ONLY: a small stand-in for the part of Boost.PropertyTree that
// DROPS' ParamCL (misc/params.h, misc/params.cpp, num/bndData.cpp) uses, so that reference translation units that
// include misc/params.h -- num/poissonsolverfactory.h, poisson/cdrdrops.cpp -- can be compiled, linked and RUN
// against the drops_b200 overlay in this image, which has no Boost.  A DROPS installation uses the real Boost;
// nothing under drops_b200/ includes this.  Written from the documented behaviour of boost::property_tree::ptree
// (string data, ordered multi-children, '.'-separated paths, stream-based value translation).
#ifndef DROPS_B200_TESTS_BOOST_MINI_PTREE_HPP
#define DROPS_B200_TESTS_BOOST_MINI_PTREE_HPP

#include <cstring>      // the real Boost headers pull this in; geom/signtraits.h relies on it (std::memcpy)
#include <iomanip>
#include <limits>
#include <list>
#include <memory>       // likewise: poisson/cdrdrops.cpp uses std::unique_ptr without including it
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>

namespace boost { namespace property_tree {

class ptree_error : public std::runtime_error {
  public:
    explicit ptree_error(const std::string& w) : std::runtime_error(w) {}
};
class ptree_bad_path : public ptree_error {
  public:
    explicit ptree_bad_path(const std::string& w) : ptree_error(w) {}
};
class ptree_bad_data : public ptree_error {
  public:
    explicit ptree_bad_data(const std::string& w) : ptree_error(w) {}
};

namespace detail {
template <typename T>
inline bool from_string(const std::string& s, T& out)
{
    std::istringstream is(s);
    is.imbue(std::locale::classic());
    is >> out;
    if (is.fail()) {                                  // bool written as true / false
        is.clear();
        is.str(s);
        is >> std::boolalpha >> out;
    }
    if (is.fail()) return false;
    is >> std::ws;
    return is.eof();
}
inline bool from_string(const std::string& s, std::string& out) { out = s; return true; }
template <typename T>
inline std::string to_string(const T& v)
{
    std::ostringstream os;
    os.imbue(std::locale::classic());
    if (std::is_floating_point<T>::value) os << std::setprecision(std::numeric_limits<T>::digits10 + 2);
    os << std::boolalpha << v;
    return os.str();
}
inline std::string to_string(const std::string& v) { return v; }
inline std::string to_string(const char* v) { return std::string(v); }
}  // namespace detail

class ptree
{
  public:
    typedef std::string key_type;
    typedef std::string data_type;
    typedef std::string path_type;
    typedef std::pair<const key_type, ptree> value_type;
    typedef std::list<value_type> container_type;
    typedef container_type::iterator iterator;
    typedef container_type::const_iterator const_iterator;
    typedef iterator assoc_iterator;
    typedef const_iterator const_assoc_iterator;
    typedef container_type::size_type size_type;

  private:
    data_type data_;
    container_type children_;

    const ptree* walk(const path_type& path) const
    {
        const ptree* node = this;
        std::string::size_type pos = 0;
        while (pos <= path.size() && !path.empty()) {
            const std::string::size_type dot = path.find('.', pos);
            const std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
            const ptree* next = nullptr;
            for (const value_type& c : node->children_)
                if (c.first == key) { next = &c.second; break; }
            if (!next) return nullptr;
            node = next;
            if (dot == std::string::npos) break;
            pos = dot + 1;
        }
        return node;
    }
    ptree& walk_create(const path_type& path)
    {
        ptree* node = this;
        std::string::size_type pos = 0;
        while (!path.empty()) {
            const std::string::size_type dot = path.find('.', pos);
            const std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
            ptree* next = nullptr;
            for (value_type& c : node->children_)
                if (c.first == key) { next = &c.second; break; }
            if (!next) { node->children_.push_back(value_type(key, ptree())); next = &node->children_.back().second; }
            node = next;
            if (dot == std::string::npos) break;
            pos = dot + 1;
        }
        return *node;
    }

  public:
    ptree() {}
    explicit ptree(const data_type& d) : data_(d) {}
    ptree(const ptree& o) : data_(o.data_), children_(o.children_) {}
    ptree& operator=(const ptree& o)            // the keys are const: rebuild the list instead of assigning nodes
    {
        if (this != &o) {
            container_type tmp(o.children_);
            data_ = o.data_;
            children_.swap(tmp);
        }
        return *this;
    }

    data_type& data() { return data_; }
    const data_type& data() const { return data_; }
    bool empty() const { return children_.empty(); }
    size_type size() const { return children_.size(); }
    iterator begin() { return children_.begin(); }
    iterator end() { return children_.end(); }
    const_iterator begin() const { return children_.begin(); }
    const_iterator end() const { return children_.end(); }
    void clear() { data_.clear(); children_.clear(); }

    iterator push_back(const value_type& v) { children_.push_back(v); return --children_.end(); }
    iterator insert(iterator where, const value_type& v) { return children_.insert(where, v); }
    template <typename It>
    void insert(iterator where, It first, It last) { children_.insert(where, first, last); }
    iterator erase(iterator where) { return children_.erase(where); }

    assoc_iterator find(const key_type& key)
    {
        for (iterator it = children_.begin(); it != children_.end(); ++it) if (it->first == key) return it;
        return children_.end();
    }
    const_assoc_iterator find(const key_type& key) const
    {
        for (const_iterator it = children_.begin(); it != children_.end(); ++it) if (it->first == key) return it;
        return children_.end();
    }
    assoc_iterator not_found() { return children_.end(); }
    const_assoc_iterator not_found() const { return children_.end(); }
    size_type count(const key_type& key) const
    {
        size_type n = 0;
        for (const value_type& c : children_) if (c.first == key) ++n;
        return n;
    }

    ptree& get_child(const path_type& path)
    {
        const ptree* p = walk(path);
        if (!p) throw ptree_bad_path("No such node (" + path + ")");
        return const_cast<ptree&>(*p);
    }
    const ptree& get_child(const path_type& path) const
    {
        const ptree* p = walk(path);
        if (!p) throw ptree_bad_path("No such node (" + path + ")");
        return *p;
    }
    ptree& put_child(const path_type& path, const ptree& child)
    {
        ptree& node = walk_create(path);
        node = child;
        return node;
    }
    ptree& add_child(const path_type& path, const ptree& child)
    {
        const std::string::size_type dot = path.rfind('.');
        ptree& parent = dot == std::string::npos ? *this : walk_create(path.substr(0, dot));
        parent.children_.push_back(value_type(dot == std::string::npos ? path : path.substr(dot + 1), child));
        return parent.children_.back().second;
    }

    template <typename T>
    T get_value() const
    {
        T out = T();
        if (!detail::from_string(data_, out))
            throw ptree_bad_data("conversion of data \"" + data_ + "\" failed");
        return out;
    }
    template <typename T>
    T get_value(const T& dflt) const
    {
        T out = T();
        return detail::from_string(data_, out) ? out : dflt;
    }
    template <typename T>
    T get(const path_type& path) const { return get_child(path).template get_value<T>(); }
    template <typename T>
    T get(const path_type& path, const T& dflt) const
    {
        const ptree* p = walk(path);
        return p ? p->template get_value<T>(dflt) : dflt;
    }
    std::string get(const path_type& path, const char* dflt) const { return get<std::string>(path, std::string(dflt)); }
    template <typename T>
    void put_value(const T& v) { data_ = detail::to_string(v); }
    template <typename T>
    ptree& put(const path_type& path, const T& v)
    {
        ptree& node = walk_create(path);
        node.put_value(v);
        return node;
    }

    bool operator==(const ptree& o) const
    {
        if (data_ != o.data_ || children_.size() != o.children_.size()) return false;
        const_iterator a = children_.begin(), b = o.children_.begin();
        for (; a != children_.end(); ++a, ++b)
            if (a->first != b->first || !(a->second == b->second)) return false;
        return true;
    }
    bool operator!=(const ptree& o) const { return !(*this == o); }
};

}}  // namespace boost::property_tree

#endif
