// qboot_b200 façade -- qboot::fs (include/qboot/my_filesystem.hpp)
#ifndef QBOOT_B200_MY_FILESYSTEM_HPP_
#define QBOOT_B200_MY_FILESYSTEM_HPP_
#include <filesystem>
namespace qboot
{
	namespace fs = std::filesystem;
}
#endif  // QBOOT_B200_MY_FILESYSTEM_HPP_
