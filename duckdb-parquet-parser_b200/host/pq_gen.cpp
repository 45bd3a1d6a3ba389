// pq_gen.cpp -- synthetic workload generator (implemented next).
