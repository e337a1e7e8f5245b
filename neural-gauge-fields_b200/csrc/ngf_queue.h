// ngf_queue.h — the 32-byte colour work item that the march kernel compacts and the colour kernel consumes.
#pragma once

namespace ngf {

struct __align__(16) QEntry {
  float c[6];                               // plane coordinates after the gauge: u_xy v_xy u_yz v_yz u_xz v_xz
  float w;                                  // compositing weight
  int id;                                   // ray (render) or row (point-wise) index; -1 = padding
};

}  // namespace ngf
