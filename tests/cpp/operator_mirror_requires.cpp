#include "filo_b200.hpp"
#include <cstdio>
int main() {
  int ok = 0;
  try { filo::PeriodicSamplesMapper p(100, 10, 50, 300, filo::InternalRangeFunction::Rate); } catch (const std::invalid_argument& e) { ok++; std::puts(e.what()); }
  try { filo::PeriodicSamplesMapper p(100, 0, 500, 300, filo::InternalRangeFunction::Rate); } catch (const std::invalid_argument& e) { ok++; std::puts(e.what()); }
  try { filo::PeriodicSamplesMapper p(100, 10, 500, std::nullopt, filo::InternalRangeFunction::Rate); } catch (const std::invalid_argument& e) { ok++; std::puts(e.what()); }
  filo::PeriodicSamplesMapper good(100, 10, 500, std::nullopt, std::nullopt);
  try { filo::FusedGpuExec ex(0); std::puts("ctx ok"); } catch (const filo::QueryError& e) { ok++; std::printf("QueryError %d %s\n", e.status, e.what()); }
  std::printf("ok=%d\n", ok);
  return 0;
}
