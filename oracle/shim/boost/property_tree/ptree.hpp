// Stand-in for the sliver of Boost.PropertyTree that visgeom's include/json.h and the TrajectoryVisualQuality
// constructor use (get_child / get<T> / get_value<T> / iteration).  TEST INFRASTRUCTURE ONLY (see oracle/shim/Eigen/Eigen):
// it lets oracle/_ref compile src/calibration/trajectory_generation.cpp where it lies; trees are built in memory by
// oracle/ref_entry.cpp, nothing is parsed.
#ifndef VISGEOM_ORACLE_PTREE_SHIM
#define VISGEOM_ORACLE_PTREE_SHIM
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace boost { namespace property_tree {

class ptree {
public:
    typedef std::pair<std::string, ptree> value_type;
    typedef std::vector<value_type>::const_iterator const_iterator;
    ptree() {}
    explicit ptree(const std::string &data) : data_(data) {}
    const ptree &get_child(const std::string &path) const
    {
        const ptree *node = this;
        size_t pos = 0;
        for (;;) {
            const size_t dot = path.find('.', pos);
            const std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
            const ptree *next = NULL;
            for (size_t i = 0; i < node->kids_.size(); i++)
                if (node->kids_[i].first == key) { next = &node->kids_[i].second; break; }
            if (!next) throw std::runtime_error("No such node (" + path + ")");
            node = next;
            if (dot == std::string::npos) return *node;
            pos = dot + 1;
        }
    }
    template <typename T> T get_value() const
    {
        std::istringstream s(data_);
        T v = T();
        s >> std::boolalpha >> v;
        return v;
    }
    template <typename T> T get(const std::string &path) const { return get_child(path).get_value<T>(); }
    const_iterator begin() const { return kids_.begin(); }
    const_iterator end() const { return kids_.end(); }
    size_t size() const { return kids_.size(); }
    // construction (used by the driver only)
    ptree &add_child(const std::string &key, const ptree &child) { kids_.push_back(value_type(key, child)); return kids_.back().second; }
    template <typename T> void put_value(const T &v) { std::ostringstream s; s.precision(17); s << v; data_ = s.str(); }
private:
    std::string data_;
    std::vector<value_type> kids_;
};

}}  // namespace boost::property_tree
#endif
