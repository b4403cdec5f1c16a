// TEST INFRASTRUCTURE — stand-in for Thirdparty/DBoW2/DBoW2/FeatureVector.h (a std::map from vocabulary node to the feature indices
// that descend through it), used only when the drop-in matcher classes are compiled without the reference tree (tests/cpp/shim_match).
#pragma once
#include <map>
#include <vector>
namespace DBoW2 {
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {
public:
    void addFeature(NodeId id, unsigned int i_feature) { (*this)[id].push_back(i_feature); }
};
}  // namespace DBoW2
