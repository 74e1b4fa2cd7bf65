// walk_tail_sim.cpp -- CPU design study (not on the product path): how a tile's shared-walk iterations distribute over the
// root's children (the work items of walk_kernel / walk_pairs_kernel, blockIdx.y) and over (root child, grandchild) pairs.
// N = 1e6 uniform sphere: 1 840 iterations per 32 targets, 80 % of them under the root child that contains the tile,
// 62 % under one grandchild.  With P ranks a rank owns N/(128 P) long items (977 at P = 8, for 1 184 resident block slots)
// and 7x as many short ones; SPH_B200_WALK_FAKE_RANKS=8 measures 1.05 ms for the rank's share against 0.77 ms ideal.
// build / run:  g++ -O2 -o /tmp/walk_tail_sim tools/walk_tail_sim.cpp && /tmp/walk_tail_sim 1000000
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <random>
#include <map>
using namespace std;
struct Node { double c[3], L; double com[3], M; int child[8]; int nch; int part; double lo[3], hi[3]; int depth; int anc1, anc2; };
vector<Node> nodes; vector<double> X, Y, Z, H;
int build(vector<int>& idx, double cx, double cy, double cz, double L, int depth, int a1, int a2) {
    int id = nodes.size(); nodes.push_back(Node());
    { Node& n = nodes[id]; n.c[0]=cx; n.c[1]=cy; n.c[2]=cz; n.L=L; n.nch=0; n.part=-1; n.depth=depth;
      n.anc1 = depth==1 ? id : a1; n.anc2 = depth==2 ? id : a2;
      n.lo[0]=cx-L; n.lo[1]=cy-L; n.lo[2]=cz-L; n.hi[0]=cx+L; n.hi[1]=cy+L; n.hi[2]=cz+L; }
    int A1 = nodes[id].anc1, A2 = nodes[id].anc2;
    if (idx.size()==1) { Node& n=nodes[id]; n.part=idx[0]; n.com[0]=X[idx[0]]; n.com[1]=Y[idx[0]]; n.com[2]=Z[idx[0]]; n.M=1; return id; }
    vector<int> sub[8];
    for (int i: idx) { int o=(X[i]-cx>0)+2*(Y[i]-cy>0)+4*(Z[i]-cz>0); sub[o].push_back(i); }
    double m=0, s[3]={0,0,0};
    for (int o=0;o<8;++o) if(!sub[o].empty()) {
        double h=L/2; int ch=build(sub[o], cx+((o&1)?h:-h), cy+((o&2)?h:-h), cz+((o&4)?h:-h), h, depth+1, A1, A2);
        Node& n=nodes[id]; n.child[n.nch++]=ch; m+=nodes[ch].M; for(int k=0;k<3;++k) s[k]+=nodes[ch].M*nodes[ch].com[k];
    }
    Node& n=nodes[id]; n.M=m; for(int k=0;k<3;++k) n.com[k]=s[k]/m; return id;
}
const double theta=0.576;
inline bool accept(int i, const Node& n){ double dx=X[i]-n.com[0],dy=Y[i]-n.com[1],dz=Z[i]-n.com[2]; double d2=dx*dx+dy*dy+dz*dz; double s=2*n.L;
    bool acc = s*s/d2<theta*theta; if(acc){ double e2=0; double p[3]={X[i],Y[i],Z[i]}; for(int k=0;k<3;++k){double a=max(max(n.lo[k]-p[k],0.0),p[k]-n.hi[k]); e2+=a*a;} acc = H[i]*H[i]/e2<0.25; } return acc; }
int main(int argc,char**argv){
    int N=argc>1?atoi(argv[1]):1000000; const int G=32;
    mt19937_64 rng(1); uniform_real_distribution<double> U(-1,1);
    while((int)X.size()<N){double x=U(rng),y=U(rng),z=U(rng); if(x*x+y*y+z*z<=1){X.push_back(x);Y.push_back(y);Z.push_back(z);}}
    double l=0; for(int i=0;i<N;++i) l=max(l,max(fabs(X[i]),max(fabs(Y[i]),fabs(Z[i]))));
    H.assign(N, 0.5*cbrt(50.0*3/(4*M_PI)/ (N/(4*M_PI/3))));
    vector<int> all(N); for(int i=0;i<N;++i) all[i]=i;
    nodes.reserve(2*N); int root=build(all,0,0,0,l,0,-1,-1);
    vector<int> order; { vector<int> st{root}; while(!st.empty()){int n=st.back(); st.pop_back(); if(nodes[n].part>=0) order.push_back(nodes[n].part); else for(int c=nodes[n].nch-1;c>=0;--c) st.push_back(nodes[n].child[c]);} }
    double s1=0,s2=0,tot=0; int ng=0;
    for (int g0=0; g0+G<=N; g0+= G*37) { ++ng;
        map<int,long> by1, by2; long total=0;
        struct E{int n; unsigned m;}; vector<E> st; st.push_back({root,0xffffffffu});
        while(!st.empty()){ E e=st.back(); st.pop_back(); Node& P=nodes[e.n];
            for(int c=0;c<P.nch;++c){ Node& n=nodes[P.child[c]]; ++total; by1[n.anc1]++; by2[n.depth>=2 ? n.anc2 : -n.anc1-2]++;
                if(n.part>=0) continue; unsigned om=0;
                for(int l2=0;l2<G;++l2) if((e.m>>l2)&1) if(!accept(order[g0+l2],n)) om|=1u<<l2;
                if(om) st.push_back({P.child[c],om}); } }
        long m1=0,m2=0; for(auto&kv:by1) m1=max(m1,kv.second); for(auto&kv:by2) m2=max(m2,kv.second);
        s1+=(double)m1/total; s2+=(double)m2/total; tot+=total; }
    printf("N=%d: iterations per group %.0f; largest root-child item %.1f%% of them; largest (root child, grandchild) item %.1f%%\n",N,tot/ng,100*s1/ng,100*s2/ng);
}
