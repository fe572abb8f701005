// TEST INFRASTRUCTURE ONLY — stand-in for Paho's <mqtt/client.h>: the reference's network/mqtt.h holds an
// mqtt::client by value; nothing here connects anywhere. The Mqtt member functions themselves are defined in
// oracle/ref_blocks_shim.cpp (publish() records the payload so tests can read what DataController produced).
#pragma once
#include <string>

namespace mqtt {
class client {
 public:
  client() {}
  client(const std::string& /*server_uri*/, const std::string& /*client_id*/) {}
};
}  // namespace mqtt
