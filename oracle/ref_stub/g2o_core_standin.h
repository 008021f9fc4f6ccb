// Stand-in for g2o's graph classes (TEST INFRASTRUCTURE, NOT PRODUCT): what base_vertex.h, base_edge.h, base_unary_edge.h(pp),
// base_binary_edge.h(pp) and the types/ sources of the reference use of optimizable_graph.h, jacobian_workspace.h, creators.h and
// the two factories — and nothing else.  Force-included (-include) in front of the reference's own translation units and of
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
  class Vertex {
   public:
    Vertex() : _id(-1), _dimension(-1), _fixed(false), _marginalized(false) {}
    virtual ~Vertex() {}
    int id() const { return _id; }
    void setId(int id) { _id = id; }
    int dimension() const { return _dimension; }
    bool fixed() const { return _fixed; }
    void setFixed(bool f) { _fixed = f; }
    bool marginalized() const { return _marginalized; }
    void setMarginalized(bool m) { _marginalized = m; }
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
    int _id, _dimension;
    bool _fixed, _marginalized;
  };
  typedef std::set<Vertex*> VertexSet;
  typedef std::vector<Vertex*> VertexContainer;

  class Edge {
   public:
    Edge() : _dimension(-1), _level(0), _robustKernel(0), _id(-1) {}
    virtual ~Edge() {}
    int id() const { return _id; }
    virtual void resize(size_t n) { _vertices.resize(n); }
    void setVertex(size_t i, Vertex* v) { _vertices[i] = v; }
    Vertex* vertex(size_t i) const { return _vertices[i]; }
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
    std::vector<Vertex*> _vertices;
    int _dimension, _level;
    RobustKernel* _robustKernel;
    int _id;
  };
};

}  // namespace g2o
#endif
