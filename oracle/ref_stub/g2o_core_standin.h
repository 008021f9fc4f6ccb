// Stand-in for g2o's graph classes (TEST INFRASTRUCTURE, NOT PRODUCT): what base_vertex.h, base_edge.h, base_unary_edge.h(pp),
// base_binary_edge.h(pp) and the types/ sources of the reference use of optimizable_graph.h, jacobian_workspace.h, creators.h and
// the two factories — and nothing else.  Vertex / Edge derive from the reference's own HyperGraph::Vertex / Edge.  Force-included (-include) in front of the reference's own translation units and of
// oracle/ref_g2o_wrap.cpp; the guards of the replaced headers are defined here so that the quote-includes inside the reference
// files find them "already seen".  Member names and meanings follow G/core/optimizable_graph.h:90-500.
#ifndef CCM_REF_STUB_G2O_CORE_STANDIN
#define CCM_REF_STUB_G2O_CORE_STANDIN
#define G2O_AIS_OPTIMIZABLE_GRAPH_HH_
#define JACOBIAN_WORKSPACE_H
#define G2O_CREATORS_H
#define G2O_FACTORY_H
#define G2O_ROBUST_KERNEL_FACTORY_H
#define G2O_REGISTER_ROBUST_KERNEL(name, classname)
#define G2O_REGISTER_TYPE(name, classname)

#include <Eigen/Core>
#include <core/hyper_graph.h>
#include <cstddef>
#include <iostream>
#include <set>
#include <vector>

namespace g2o {

class RobustKernel;

class JacobianWorkspace {  // G/core/jacobian_workspace.h:69: one scratch block per vertex slot of the edge
 public:
  JacobianWorkspace() { slot_[0].resize(64, 0.); slot_[1].resize(64, 0.); }
  double* workspaceForVertex(int i) { return slot_[i].data(); }
 private:
  std::vector<double> slot_[2];
};

class OptimizableGraph {
 public:
  // G/core/optimizable_graph.h:90-500, on top of the reference's own HyperGraph::Vertex / Edge (G/core/hyper_graph.h, compiled in place)
  class Vertex : public HyperGraph::Vertex {
   public:
    Vertex() : HyperGraph::Vertex(-1), _hessianIndex(-1), _dimension(-1), _colInHessian(-1), _fixed(false), _marginalized(false) {}
    virtual ~Vertex() {}
    int dimension() const { return _dimension; }
    bool fixed() const { return _fixed; }
    void setFixed(bool f) { _fixed = f; }
    bool marginalized() const { return _marginalized; }
    void setMarginalized(bool m) { _marginalized = m; }
    int hessianIndex() const { return _hessianIndex; }
    void setHessianIndex(int ti) { _hessianIndex = ti; }
    int colInHessian() const { return _colInHessian; }
    void setColInHessian(int c) { _colInHessian = c; }
    void oplus(const double* v) { oplusImpl(v); updateCache(); }  // G/core/optimizable_graph.h:259-263
    void setToOrigin() { setToOriginImpl(); updateCache(); }
    void updateCache() {}
    virtual void push() = 0;
    virtual void pop() = 0;
    virtual void discardTop() = 0;
    virtual const double& hessian(int i, int j) const = 0;
    virtual void mapHessianMemory(double* d) = 0;
    virtual void clearQuadraticForm() = 0;
    virtual int copyB(double* b) const = 0;
    virtual bool read(std::istream& is) = 0;
    virtual bool write(std::ostream& os) const = 0;
   protected:
    virtual void oplusImpl(const double* v) = 0;
    virtual void setToOriginImpl() = 0;
    int _hessianIndex, _dimension, _colInHessian;
    bool _fixed, _marginalized;
  };
  typedef HyperGraph::VertexSet VertexSet;
  typedef std::vector<Vertex*> VertexContainer;

  class Edge : public HyperGraph::Edge {
   public:
    Edge() : HyperGraph::Edge(-1), _dimension(-1), _level(0), _robustKernel(0) {}
    virtual ~Edge() {}
    RobustKernel* robustKernel() const { return _robustKernel; }
    void setRobustKernel(RobustKernel* k) { _robustKernel = k; }  // not owned here
    int level() const { return _level; }
    void setLevel(int l) { _level = l; }
    int dimension() const { return _dimension; }
    virtual void computeError() = 0;
    virtual double chi2() const = 0;
    virtual void constructQuadraticForm() = 0;
    virtual void linearizeOplus(JacobianWorkspace& w) = 0;
    virtual void mapHessianMemory(double* d, int i, int j, bool rowMajor) = 0;
    virtual bool allVerticesFixed() const = 0;
    virtual bool read(std::istream& is) = 0;
    virtual bool write(std::ostream& os) const = 0;
    virtual double initialEstimatePossible(const VertexSet&, Vertex*) { return -1.; }
    virtual void initialEstimate(const VertexSet&, Vertex*) = 0;
    virtual Vertex* createFrom() { return 0; }
    virtual Vertex* createTo() { return 0; }
   protected:
    int _dimension, _level;
    RobustKernel* _robustKernel;
  };
  typedef std::vector<Edge*> EdgeContainer;
};

}  // namespace g2o
#endif
